"""Python sqlite3 restatement of the reference's on-disk format -- TEST INFRASTRUCTURE.

/root/reference/cpp/database.cc:77-89 (pragmas), :108-135 (tables), :137-158 (raw little-endian
blobs: keypoints rows x 2 f32; src_keypoints_indices u32; tgt_keypoints rows x 2 f32;
flow_errors f32)."""
from __future__ import annotations

import sqlite3

import numpy as np

SCHEMA = [
    """CREATE TABLE IF NOT EXISTS keypoints(
            image_id   INTEGER  PRIMARY KEY  NOT NULL,
            rows       INTEGER               NOT NULL,
            keypoints  BLOB                  NOT NULL);""",
    """CREATE TABLE IF NOT EXISTS optical_flow(
            image_id_from           INTEGER  NOT NULL,
            image_id_to             INTEGER  NOT NULL,
            rows                    INTEGER  NOT NULL,
            src_keypoints_indices   BLOB     NOT NULL,
            tgt_keypoints           BLOB     NOT NULL,
            flow_errors             BLOB     NOT NULL,
            PRIMARY KEY(image_id_from, image_id_to),
            FOREIGN KEY(image_id_from) REFERENCES keypoints(image_id) ON DELETE CASCADE);""",
]


class Database:
    def __init__(self, path: str):
        self.con = sqlite3.connect(path, isolation_level=None)
        for p in ("synchronous=OFF", "journal_mode=WAL", "temp_store=MEMORY", "foreign_keys=ON", "auto_vacuum=1"):
            self.con.execute("PRAGMA " + p)
        for s in SCHEMA:
            self.con.execute(s)

    def close(self):
        self.con.close()

    def write_keypoints(self, image_id: int, kps: np.ndarray):
        kps = np.ascontiguousarray(kps, "<f4").reshape(-1, 2)
        self.con.execute("INSERT INTO keypoints(image_id, rows, keypoints) VALUES(?, ?, ?);",
                         (image_id, len(kps), kps.tobytes()))

    def read_keypoints(self, image_id: int) -> np.ndarray:
        r = self.con.execute("SELECT rows, keypoints FROM keypoints WHERE image_id = ?;", (image_id,)).fetchone()
        if r is None:
            return np.zeros((0, 2), np.float32)
        return np.frombuffer(r[1], "<f4").reshape(r[0], 2).copy()

    def write_image_pair_flow(self, a: int, b: int, idx, tgt, err):
        idx = np.ascontiguousarray(idx, "<u4")
        tgt = np.ascontiguousarray(tgt, "<f4").reshape(-1, 2)
        err = np.ascontiguousarray(err, "<f4")
        assert len(idx) == len(tgt) == len(err)
        self.con.execute("INSERT INTO optical_flow(image_id_from, image_id_to, rows, src_keypoints_indices, "
                         "tgt_keypoints, flow_errors) VALUES(?, ?, ?, ?, ?, ?);",
                         (a, b, len(idx), idx.tobytes(), tgt.tobytes(), err.tobytes()))

    def read_image_pair_flow(self, a: int, b: int):
        r = self.con.execute("SELECT rows, src_keypoints_indices, tgt_keypoints, flow_errors FROM optical_flow "
                             "WHERE image_id_from = ? AND image_id_to = ?;", (a, b)).fetchone()
        if r is None:
            return None
        return (np.frombuffer(r[1], "<u4").copy(), np.frombuffer(r[2], "<f4").reshape(r[0], 2).copy(),
                np.frombuffer(r[3], "<f4").copy())

    def find_optical_flows_to_image(self, b: int):
        return [r[0] for r in self.con.execute("SELECT image_id_from FROM optical_flow WHERE image_id_to = ?", (b,))]

    def find_optical_flows_from_image(self, a: int):
        return [r[0] for r in self.con.execute("SELECT image_id_to FROM optical_flow WHERE image_id_from = ?", (a,))]

    def pairs(self):
        return [tuple(r) for r in self.con.execute("SELECT image_id_from, image_id_to FROM optical_flow ORDER BY 1, 2")]

    def frames(self):
        return [r[0] for r in self.con.execute("SELECT image_id FROM keypoints ORDER BY 1")]
