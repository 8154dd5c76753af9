// chrom/result.hpp — the reference's `Result<T, String>` for a C++ host.
// chrom-rs reports every failure on the solver boundary as `Err(String)` (src/solver/traits.rs:624-628);
// the C++ mirror keeps that shape so call sites and tests read like the reference's.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>

namespace chrom {

struct Unit {};

struct ErrTag {
  std::string message;
};
inline ErrTag Err(std::string message) { return ErrTag{std::move(message)}; }

template <class T>
class Result {
 public:
  Result(T value) : v_(std::in_place_index<0>, std::move(value)) {}
  Result(ErrTag e) : v_(std::in_place_index<1>, std::move(e.message)) {}

  bool is_ok() const { return v_.index() == 0; }
  bool is_err() const { return v_.index() == 1; }
  explicit operator bool() const { return is_ok(); }

  // `unwrap` / `expect` panic in Rust; here they throw std::runtime_error carrying the Err text.
  T& unwrap() & {
    if (is_err()) throw std::runtime_error("called `Result::unwrap()` on an `Err` value: \"" + std::get<1>(v_) + "\"");
    return std::get<0>(v_);
  }
  const T& unwrap() const& {
    if (is_err()) throw std::runtime_error("called `Result::unwrap()` on an `Err` value: \"" + std::get<1>(v_) + "\"");
    return std::get<0>(v_);
  }
  T unwrap() && {
    if (is_err()) throw std::runtime_error("called `Result::unwrap()` on an `Err` value: \"" + std::get<1>(v_) + "\"");
    return std::move(std::get<0>(v_));
  }
  T expect(const std::string& what) && {
    if (is_err()) throw std::runtime_error(what + ": \"" + std::get<1>(v_) + "\"");
    return std::move(std::get<0>(v_));
  }
  const std::string& unwrap_err() const {
    if (is_ok()) throw std::runtime_error("called `Result::unwrap_err()` on an `Ok` value");
    return std::get<1>(v_);
  }

 private:
  std::variant<T, std::string> v_;
};

inline Result<Unit> Ok() { return Result<Unit>(Unit{}); }

}  // namespace chrom
