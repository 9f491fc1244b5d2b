// Stand-in for jansson: caffe/util/pb2json.h (included by the reference's solver.cpp) only names the type.
#pragma once
typedef struct json_t json_t;
