#pragma once
#include <cstddef>
typedef struct MDB_env MDB_env; typedef struct MDB_txn MDB_txn; typedef struct MDB_cursor MDB_cursor; typedef unsigned int MDB_dbi;
typedef struct MDB_val { size_t mv_size; void* mv_data; } MDB_val;
