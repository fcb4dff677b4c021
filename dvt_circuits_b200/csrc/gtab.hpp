// handle of the fixed-base table of the G1 generator (layout, digit logic and builder: feldman.cuh)
#pragma once
#include <cstdint>
namespace dkgv {
struct GTab {
  const uint32_t* p;
  uint32_t bits, windows;
};
}  // namespace dkgv
