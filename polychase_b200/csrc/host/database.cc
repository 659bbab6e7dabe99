#include "database.h"

#include <cstdlib>

#include <cstring>
#include <stdexcept>

namespace pch {

int Database::Call(int rc, int line) const {
    switch (rc) {
        case SQLITE_OK:
        case SQLITE_ROW:
        case SQLITE_DONE:
            return rc;
        default:      // database.cc:27-38: SQLite failures are std::runtime_error
            throw std::runtime_error(std::string("SQLite error [") + __FILE__ + ":" + std::to_string(line) + "]: " +
                                     sqlite3_errstr(rc) + (db_ ? std::string(" (") + sqlite3_errmsg(db_) + ")" : ""));
    }
}
#define SQL_CALL(x) Call((x), __LINE__)

void Database::Exec(const char* sql) const {
    char* err = nullptr;
    const int rc = sqlite3_exec(db_, sql, nullptr, nullptr, &err);
    if (rc != SQLITE_OK) {
        const std::string msg = std::string("SQLite error [") + __FILE__ + "]: " + (err ? err : "Unknown error");
        sqlite3_free(err);
        throw std::runtime_error(msg);
    }
}

sqlite3_stmt* Database::Prepare(const char* sql) {
    sqlite3_stmt* st = nullptr;
    SQL_CALL(sqlite3_prepare_v2(db_, sql, -1, &st, nullptr));
    return st;
}

Database::Database(const std::string& path) { Open(path); }
Database::~Database() { Close(); }

void Database::Open(const std::string& path) {
    Close();
    SQL_CALL(sqlite3_open_v2(path.c_str(), &db_, SQLITE_OPEN_READWRITE | SQLITE_OPEN_CREATE | SQLITE_OPEN_NOMUTEX,
                             nullptr));
    // A database created here gets 64 KB pages (no effect on an existing file; any SQLite build, the reference's
    // included, reads either): a flow row carries ~136 KB of blobs at 8 000 features, which 4 KB pages spread over
    // 34 overflow pages and WAL frames each -- 6 300 instead of 3 400 pair rows/s on the write path that bounds the
    // analyze pass of the drop-in.  PC_DB_PAGE_SIZE=0 keeps SQLite's default.
    {
        const char* e = getenv("PC_DB_PAGE_SIZE");
        const int page = e ? atoi(e) : 65536;
        if (page >= 512) Exec(("PRAGMA page_size=" + std::to_string(page)).c_str());
    }
    // database.cc:77-89
    Exec("PRAGMA synchronous=OFF");
    Exec("PRAGMA journal_mode=WAL");
    Exec("PRAGMA temp_store=MEMORY");
    Exec("PRAGMA foreign_keys=ON");
    Exec("PRAGMA auto_vacuum=1");
    // database.cc:108-135
    Exec("CREATE TABLE IF NOT EXISTS keypoints("
         "    image_id   INTEGER  PRIMARY KEY  NOT NULL,"
         "    rows       INTEGER               NOT NULL,"
         "    keypoints  BLOB                  NOT NULL);");
    Exec("CREATE TABLE IF NOT EXISTS optical_flow("
         "    image_id_from           INTEGER  NOT NULL,"
         "    image_id_to             INTEGER  NOT NULL,"
         "    rows                    INTEGER  NOT NULL,"
         "    src_keypoints_indices   BLOB     NOT NULL,"
         "    tgt_keypoints           BLOB     NOT NULL,"
         "    flow_errors             BLOB     NOT NULL,"
         "    PRIMARY KEY(image_id_from, image_id_to),"
         "    FOREIGN KEY(image_id_from) REFERENCES keypoints(image_id) ON DELETE CASCADE);");
    read_kps_ = Prepare("SELECT rows, keypoints FROM keypoints WHERE image_id = ?;");
    write_kps_ = Prepare("INSERT INTO keypoints(image_id, rows, keypoints) VALUES(?, ?, ?);");
    read_flow_ = Prepare("SELECT rows, src_keypoints_indices, tgt_keypoints, flow_errors FROM optical_flow WHERE "
                         "image_id_from = ? AND image_id_to = ?;");
    write_flow_ = Prepare("INSERT INTO optical_flow(image_id_from, image_id_to, rows, src_keypoints_indices, "
                          "tgt_keypoints, flow_errors) VALUES(?, ?, ?, ?, ?, ?);");
    flows_from_ = Prepare("SELECT image_id_to FROM optical_flow WHERE image_id_from = ?");
    flows_to_ = Prepare("SELECT image_id_from FROM optical_flow WHERE image_id_to = ?");
    kps_exist_ = Prepare("SELECT 1 FROM keypoints WHERE image_id = ?;");
    flow_exist_ = Prepare("SELECT 1 FROM optical_flow WHERE image_id_from = ? AND image_id_to = ?;");
    min_id_ = Prepare("SELECT MIN(image_id) FROM keypoints;");
    max_id_ = Prepare("SELECT MAX(image_id) FROM keypoints;");
}

void Database::Close() {
    if (!db_) return;
    for (sqlite3_stmt** s : {&read_kps_, &write_kps_, &read_flow_, &write_flow_, &flows_from_, &flows_to_, &kps_exist_,
                             &flow_exist_, &min_id_, &max_id_}) {
        if (*s) sqlite3_finalize(*s);
        *s = nullptr;
    }
    sqlite3_close_v2(db_);
    db_ = nullptr;
}

void Database::Begin() { Exec("BEGIN"); }
void Database::Commit() { Exec("COMMIT"); }

template <typename T>
static void ReadBlob(sqlite3_stmt* st, int rows, int col, std::vector<T>& vec) {
    vec.clear();
    vec.resize(rows);
    const size_t bytes = static_cast<size_t>(sqlite3_column_bytes(st, col));
    PCH_CHECK(vec.size() * sizeof(T) == bytes);                      // database.cc:145
    if (bytes) memcpy(reinterpret_cast<char*>(vec.data()), sqlite3_column_blob(st, col), bytes);
}

void Database::ReadKeypoints(int32_t image_id, Keypoints& keypoints) const {
    sqlite3_stmt* st = read_kps_;
    SQL_CALL(sqlite3_bind_int(st, 1, image_id));
    const int rc = SQL_CALL(sqlite3_step(st));
    if (rc != SQLITE_ROW) {                                          // database.cc:168-171: leaves the vector alone
        SQL_CALL(sqlite3_reset(st));
        return;
    }
    const int rows = sqlite3_column_int(st, 0);
    PCH_CHECK(rows >= 0);
    ReadBlob(st, rows, 1, keypoints);
    SQL_CALL(sqlite3_reset(st));
}

Keypoints Database::ReadKeypoints(int32_t image_id) const {
    Keypoints k;
    ReadKeypoints(image_id, k);
    return k;
}

void Database::WriteKeypoints(int32_t image_id, const float* xy, size_t rows) {
    sqlite3_stmt* st = write_kps_;
    static const char kEmpty = 0;
    SQL_CALL(sqlite3_bind_int(st, 1, image_id));
    SQL_CALL(sqlite3_bind_int(st, 2, (int)rows));
    SQL_CALL(sqlite3_bind_blob(st, 3, rows ? (const void*)xy : (const void*)&kEmpty, (int)(rows * 2 * sizeof(float)),
                               SQLITE_STATIC));
    const int rc = sqlite3_step(st);
    sqlite3_reset(st);
    SQL_CALL(rc);
}

void Database::WriteKeypoints(int32_t image_id, const Keypoints& keypoints) {
    WriteKeypoints(image_id, keypoints.empty() ? nullptr : keypoints[0].data(), keypoints.size());
}

void Database::WriteImagePairFlow(int32_t from, int32_t to, const uint32_t* idx, const float* tgt_xy, const float* err,
                                  size_t rows) {
    sqlite3_stmt* st = write_flow_;
    static const char kEmpty = 0;
    SQL_CALL(sqlite3_bind_int(st, 1, from));
    SQL_CALL(sqlite3_bind_int(st, 2, to));
    SQL_CALL(sqlite3_bind_int(st, 3, (int)rows));
    SQL_CALL(sqlite3_bind_blob(st, 4, rows ? (const void*)idx : &kEmpty, (int)(rows * sizeof(uint32_t)), SQLITE_STATIC));
    SQL_CALL(sqlite3_bind_blob(st, 5, rows ? (const void*)tgt_xy : &kEmpty, (int)(rows * 2 * sizeof(float)), SQLITE_STATIC));
    SQL_CALL(sqlite3_bind_blob(st, 6, rows ? (const void*)err : &kEmpty, (int)(rows * sizeof(float)), SQLITE_STATIC));
    const int rc = sqlite3_step(st);
    sqlite3_reset(st);
    SQL_CALL(rc);
}

void Database::WriteImagePairFlow(int32_t from, int32_t to, const KeypointsIndices& idx, const Keypoints& tgt,
                                  const FlowErrors& err) {
    const size_t rows = idx.size();
    PCH_CHECK(tgt.size() == rows);                                   // database.cc:202-203
    PCH_CHECK(err.size() == rows);
    WriteImagePairFlow(from, to, idx.data(), rows ? tgt[0].data() : nullptr, err.data(), rows);
}

void Database::WriteImagePairFlow(const ImagePairFlow& f) {
    WriteImagePairFlow(f.image_id_from, f.image_id_to, f.src_kps_indices, f.tgt_kps, f.flow_errors);
}

void Database::ReadImagePairFlow(int32_t from, int32_t to, ImagePairFlow& flow) const {
    sqlite3_stmt* st = read_flow_;
    SQL_CALL(sqlite3_bind_int(st, 1, from));
    SQL_CALL(sqlite3_bind_int(st, 2, to));
    const int rc = SQL_CALL(sqlite3_step(st));
    if (rc != SQLITE_ROW) {
        SQL_CALL(sqlite3_reset(st));
        return;
    }
    const int rows = sqlite3_column_int(st, 0);
    ReadBlob(st, rows, 1, flow.src_kps_indices);
    ReadBlob(st, rows, 2, flow.tgt_kps);
    ReadBlob(st, rows, 3, flow.flow_errors);
    flow.image_id_from = from;
    flow.image_id_to = to;
    SQL_CALL(sqlite3_reset(st));
}

ImagePairFlow Database::ReadImagePairFlow(int32_t from, int32_t to) const {
    ImagePairFlow f;
    ReadImagePairFlow(from, to, f);
    return f;
}

std::vector<int32_t> Database::FindOpticalFlowsFromImage(int32_t from) const {
    std::vector<int32_t> out;
    sqlite3_stmt* st = flows_from_;
    SQL_CALL(sqlite3_bind_int(st, 1, from));
    while (SQL_CALL(sqlite3_step(st)) == SQLITE_ROW) out.push_back(sqlite3_column_int(st, 0));
    SQL_CALL(sqlite3_reset(st));
    return out;
}

std::vector<int32_t> Database::FindOpticalFlowsToImage(int32_t to) const {
    std::vector<int32_t> out;
    sqlite3_stmt* st = flows_to_;
    SQL_CALL(sqlite3_bind_int(st, 1, to));
    while (SQL_CALL(sqlite3_step(st)) == SQLITE_ROW) out.push_back(sqlite3_column_int(st, 0));
    SQL_CALL(sqlite3_reset(st));
    return out;
}

bool Database::KeypointsExist(int32_t image_id) const {
    sqlite3_stmt* st = kps_exist_;
    SQL_CALL(sqlite3_bind_int(st, 1, image_id));
    const bool e = SQL_CALL(sqlite3_step(st)) == SQLITE_ROW;
    SQL_CALL(sqlite3_reset(st));
    return e;
}

bool Database::ImagePairFlowExists(int32_t from, int32_t to) const {
    sqlite3_stmt* st = flow_exist_;
    SQL_CALL(sqlite3_bind_int(st, 1, from));
    SQL_CALL(sqlite3_bind_int(st, 2, to));
    const bool e = SQL_CALL(sqlite3_step(st)) == SQLITE_ROW;
    SQL_CALL(sqlite3_reset(st));
    return e;
}

int32_t Database::GetMinImageIdWithKeypoints() const {
    sqlite3_stmt* st = min_id_;
    const int rc = SQL_CALL(sqlite3_step(st));
    int32_t id = kInvalidId;
    if (rc == SQLITE_ROW) id = sqlite3_column_int(st, 0);
    SQL_CALL(sqlite3_reset(st));
    return id;
}

int32_t Database::GetMaxImageIdWithKeypoints() const {
    sqlite3_stmt* st = max_id_;
    const int rc = SQL_CALL(sqlite3_step(st));
    int32_t id = kInvalidId;
    if (rc == SQLITE_ROW) id = sqlite3_column_int(st, 0);
    SQL_CALL(sqlite3_reset(st));
    return id;
}

}  // namespace pch
