// TEST INFRASTRUCTURE -- stand-in for gar/riccati-base.hpp:13-37 (the same six virtuals, same signatures).
#pragma once
#include <optional>
#include "aligator/math.hpp"
namespace aligator {
namespace gar {
template <typename Scalar> struct LqrKnotTpl;
template <typename _Scalar> class RiccatiSolverBase {
public:
  using Scalar = _Scalar;
  using LqrKnot = LqrKnotTpl<Scalar>;
  ALIGATOR_DYNAMIC_TYPEDEFS_WITH_ROW_TYPES(Scalar);
  virtual bool backward(const Scalar mueq) = 0;
  virtual bool forward(std::vector<VectorXs> &xs, std::vector<VectorXs> &us, std::vector<VectorXs> &vs,
                       std::vector<VectorXs> &lbdas, const std::optional<ConstVectorRef> &theta_ = std::nullopt) const = 0;
  virtual void cycleAppend(const LqrKnot &knot) = 0;
  virtual void collapseFeedback() {}
  virtual VectorRef getFeedforward(size_t) = 0;
  virtual RowMatrixRef getFeedback(size_t) = 0;
  virtual ~RiccatiSolverBase() = default;
};
} // namespace gar
} // namespace aligator
