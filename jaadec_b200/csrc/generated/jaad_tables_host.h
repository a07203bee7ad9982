// Default instantiation of the generated table header: every float table
// becomes `static const uint32_t NAME_bits[N]` (IEEE-754 binary32 patterns) plus
// the accessor macro `JT(NAME)` yielding `const float*`; integer tables become
// plain static arrays.  Shared by the oracle, the bitstream generator and the
// engine's host side (which uploads the same bit patterns to the device).
// Internal linkage on purpose: each translation unit that includes this gets
// its own read-only copy, so no link-order coupling between the libraries.
#pragma once
#include <cstdint>

namespace jaad_tables {

#define JAAD_TABLE_F32(name, n, dims)                                      \
  static constexpr int name##_N = n;                                       \
  alignas(16) static const uint32_t name##_bits[n] = {
#define JAAD_TABLE_INT(name, type, n)                                      \
  static constexpr int name##_N = n;                                       \
  static const type name[n] = {
#define JAAD_TABLE_END };

#include "jaad_tables.h"

#undef JAAD_TABLE_F32
#undef JAAD_TABLE_INT
#undef JAAD_TABLE_END

}  // namespace jaad_tables

#define JT(name) (reinterpret_cast<const float*>(::jaad_tables::name##_bits))
