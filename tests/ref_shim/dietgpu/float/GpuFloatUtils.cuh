// shim (test support only): the reference's FloatTest uses FloatTypeInfo<FT>::WordT to pick the
// host word type of a float kind; nothing else of the reference's device-side utilities.
#pragma once
#include <cstdint>

#include "dietgpu_b200_compat.hpp"

namespace dietgpu {
template <FloatType FT>
struct FloatTypeInfo;
template <>
struct FloatTypeInfo<FloatType::kFloat16> {
  using WordT = uint16_t;
};
template <>
struct FloatTypeInfo<FloatType::kBFloat16> {
  using WordT = uint16_t;
};
template <>
struct FloatTypeInfo<FloatType::kFloat32> {
  using WordT = uint32_t;
};
}  // namespace dietgpu
