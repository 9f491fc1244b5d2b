// Stand-in: caffe/util/pb2json.h (included by the reference's solver.cpp) only names these types in declarations.
#pragma once
#include "google/protobuf/message.h"
namespace google { namespace protobuf { class Reflection; class FieldDescriptor; } }
