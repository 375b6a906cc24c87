// nvblox/integrators/weighting_function.h (reference: :11-18)
#pragma once
namespace nvblox {
enum class WeightingFunctionType {
  kConstantWeight,
  kConstantDropoffWeight,
  kInverseSquareWeight,
  kInverseSquareDropoffWeight,
  kInverseSquareTsdfDistancePenalty,
  kLinearWithMax
};
}  // namespace nvblox
