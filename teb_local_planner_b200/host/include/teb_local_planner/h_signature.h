/*
 * h_signature.h — equivalence classes of candidate trajectories (which homotopy class a band belongs to).
 *
 * Mirrors teb_local_planner::EquivalenceClass (include/teb_local_planner/equivalence_relations.h:52-96), HSignature
 * (include/teb_local_planner/h_signature.h:67-240) and HSignature3d (:252-424): isEqual / isValid / isReasonable keep the
 * reference's rules. The VALUES are not computed here: HomotopyClassPlanner computes them for all its candidates in one
 * batched device call (tebgpu_h_signature, csrc/teb_hsig.cuh) and stores them in these objects.
 */
#ifndef TEB_B200_H_SIGNATURE_H_
#define TEB_B200_H_SIGNATURE_H_

#include <cmath>
#include <complex>
#include <memory>
#include <vector>

#include "teb_local_planner/teb_config.h"

namespace teb_local_planner {

class EquivalenceClass {
 public:
  virtual ~EquivalenceClass() {}
  virtual bool isEqual(const EquivalenceClass& other) const = 0;
  virtual bool isValid() const = 0;
  virtual bool isReasonable() const = 0;
};
typedef std::shared_ptr<EquivalenceClass> EquivalenceClassPtr;
typedef std::vector<std::pair<EquivalenceClassPtr, bool>> EquivalenceClassContainer; /* class, locked (:103) */

class HSignature : public EquivalenceClass {
 public:
  HSignature(const TebConfig& cfg, const std::complex<double>& value) : cfg_(&cfg), hsignature_(value) {}
  bool isEqual(const EquivalenceClass& other) const override { /* h_signature.h:191-207 */
    const HSignature* o = dynamic_cast<const HSignature*>(&other);
    if (!o) return false;
    return std::fabs(o->hsignature_.real() - hsignature_.real()) <= cfg_->hcp.h_signature_threshold &&
           std::fabs(o->hsignature_.imag() - hsignature_.imag()) <= cfg_->hcp.h_signature_threshold;
  }
  bool isValid() const override { return std::isfinite(hsignature_.real()) && std::isfinite(hsignature_.imag()); }
  bool isReasonable() const override { return true; }
  const std::complex<double>& value() const { return hsignature_; }

 private:
  const TebConfig* cfg_;
  std::complex<double> hsignature_;
};

class HSignature3d : public EquivalenceClass {
 public:
  HSignature3d(const TebConfig& cfg, const std::vector<double>& values) : cfg_(&cfg), hsignature3d_(values) {}
  bool isEqual(const EquivalenceClass& other) const override { /* h_signature.h:366-388 */
    const HSignature3d* o = dynamic_cast<const HSignature3d*>(&other);
    if (!o || o->hsignature3d_.size() != hsignature3d_.size()) return false;
    for (size_t i = 0; i < hsignature3d_.size(); ++i) {
      /* an obstacle far from either trajectory does not take part */
      if (std::fabs(o->hsignature3d_[i]) < cfg_->hcp.h_signature_threshold || std::fabs(hsignature3d_[i]) < cfg_->hcp.h_signature_threshold)
        continue;
      if (sign(o->hsignature3d_[i]) != sign(hsignature3d_[i])) return false;
    }
    return true;
  }
  bool isValid() const override {
    for (double v : hsignature3d_)
      if (!std::isfinite(v)) return false;
    return true;
  }
  bool isReasonable() const override { /* a value above 1 means a loop around that obstacle (:406-414) */
    for (double v : hsignature3d_)
      if (v > 1.0) return false;
    return true;
  }
  const std::vector<double>& values() const { return hsignature3d_; }

 private:
  static int sign(double v) { return (v > 0) - (v < 0); }
  const TebConfig* cfg_;
  std::vector<double> hsignature3d_;
};

}  // namespace teb_local_planner
#endif
