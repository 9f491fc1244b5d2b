#pragma once
#include <string>
namespace google { namespace protobuf { class Message { public: virtual ~Message() {} }; } }
