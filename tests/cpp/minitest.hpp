// Minimal test harness for the C++ host layer: TEST(name) { CHECK(...); } and a main() that runs
// every registered test (or those whose name contains argv[1]) and exits non-zero on failure.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace minitest {
struct Case {
  const char* name;
  std::function<void()> fn;
};
inline std::vector<Case>& registry() {
  static std::vector<Case> r;
  return r;
}
struct Registrar {
  Registrar(const char* name, std::function<void()> fn) { registry().push_back({name, std::move(fn)}); }
};
struct Failure : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void fail(const char* file, int line, const std::string& what) {
  throw Failure(std::string(file) + ":" + std::to_string(line) + ": " + what);
}
inline int run(int argc, char** argv) {
  int failed = 0, ran = 0;
  for (auto& c : registry()) {
    if (argc > 1 && !std::strstr(c.name, argv[1])) continue;
    ++ran;
    try {
      c.fn();
      std::printf("ok      %s\n", c.name);
    } catch (const std::exception& e) {
      ++failed;
      std::printf("FAILED  %s\n        %s\n", c.name, e.what());
    }
  }
  std::printf("%d run, %d failed\n", ran, failed);
  return failed ? 1 : 0;
}
}  // namespace minitest

#define TEST(name)                                                    \
  static void name();                                                 \
  static minitest::Registrar reg_##name(#name, name);                 \
  static void name()
#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) minitest::fail(__FILE__, __LINE__, "CHECK(" #cond ")"); \
  } while (0)
#define CHECK_MSG(cond, msg)                                          \
  do {                                                                \
    if (!(cond)) minitest::fail(__FILE__, __LINE__, std::string("CHECK(" #cond "): ") + (msg)); \
  } while (0)
#define CHECK_EQ(a, b)                                                \
  do {                                                                \
    if (!((a) == (b))) minitest::fail(__FILE__, __LINE__, "CHECK_EQ(" #a ", " #b ")"); \
  } while (0)
#define CHECK_STR_EQ(a, b)                                            \
  do {                                                                \
    const std::string _a = (a), _b = (b);                             \
    if (_a != _b) minitest::fail(__FILE__, __LINE__, "\"" + _a + "\" != \"" + _b + "\""); \
  } while (0)
#define CHECK_NEAR(a, b, tol)                                         \
  do {                                                                \
    const double _a = (a), _b = (b);                                  \
    if (!(std::fabs(_a - _b) <= (tol)))                               \
      minitest::fail(__FILE__, __LINE__, "|" + std::to_string(_a) + " - " + std::to_string(_b) + "| > " #tol); \
  } while (0)
#define CHECK_THROWS(expr)                                            \
  do {                                                                \
    bool _threw = false;                                              \
    try {                                                             \
      (void)(expr);                                                   \
    } catch (const minitest::Failure&) {                              \
      throw;                                                          \
    } catch (const std::exception&) {                                 \
      _threw = true;                                                  \
    }                                                                 \
    if (!_threw) minitest::fail(__FILE__, __LINE__, "expected a throw: " #expr); \
  } while (0)
