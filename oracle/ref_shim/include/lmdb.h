#pragma once
// Shim of the LMDB C API surface the reference's data layers use.  The functions are implemented by
// oracle/ref_shim/ref_driver.cpp as an in-memory record list (test infrastructure: lets the reference's own
// VideoSampledShotsDataLayer run unmodified over a synthetic dataset).
#include <cstddef>
typedef struct MDB_env MDB_env; typedef struct MDB_txn MDB_txn; typedef struct MDB_cursor MDB_cursor; typedef unsigned int MDB_dbi;
typedef struct MDB_val { size_t mv_size; void* mv_data; } MDB_val;
typedef int mdb_mode_t;
enum { MDB_SUCCESS = 0, MDB_NOTFOUND = -30798 };
enum { MDB_RDONLY = 0x20000, MDB_NOTLS = 0x200000 };
typedef enum MDB_cursor_op { MDB_FIRST = 0, MDB_GET_CURRENT = 4, MDB_NEXT = 8 } MDB_cursor_op;
#define VV_SHIM_LIVE_OBJECT (-0x5EED)
extern "C" {
int mdb_env_create(MDB_env** env);
int mdb_env_set_mapsize(MDB_env* env, size_t size);
int mdb_env_open(MDB_env* env, const char* path, unsigned int flags, mdb_mode_t mode);
int mdb_txn_begin(MDB_env* env, MDB_txn* parent, unsigned int flags, MDB_txn** txn);
int mdb_open(MDB_txn* txn, const char* name, unsigned int flags, MDB_dbi* dbi);
int mdb_cursor_open(MDB_txn* txn, MDB_dbi dbi, MDB_cursor** cursor);
int mdb_cursor_get(MDB_cursor* cursor, MDB_val* key, MDB_val* data, MDB_cursor_op op);
void mdb_cursor_close(MDB_cursor* cursor);
void mdb_close(MDB_env* env, MDB_dbi dbi);
void mdb_txn_abort(MDB_txn* txn);
void mdb_env_close(MDB_env* env);
}
