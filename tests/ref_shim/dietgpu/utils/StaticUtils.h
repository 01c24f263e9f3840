// shim (test support only): integer helpers the reference's tests use.
#pragma once
namespace dietgpu {
template <typename U, typename V>
constexpr auto divUp(U a, V b) -> decltype(a + b) {
  return (a + b - 1) / b;
}
template <typename U, typename V>
constexpr auto roundUp(U a, V b) -> decltype(a + b) {
  return divUp(a, b) * b;
}
template <typename U, typename V>
constexpr auto roundDown(U a, V b) -> decltype(a + b) {
  return (a / b) * b;
}
}  // namespace dietgpu
