#pragma once
#include <random>
namespace boost { typedef std::mt19937 mt19937; namespace random { typedef std::mt19937 mt19937; } }
