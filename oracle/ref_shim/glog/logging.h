// Minimal stand-in for <glog/logging.h> (glog is not installed and cannot be fetched offline).
// TEST INFRASTRUCTURE ONLY: used solely to compile the UNMODIFIED reference sources from
// /root/reference into oracle/_ref/libggnn_ref.so (see oracle/build_ref.sh). Not part of the product.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace ggnn_shim_log {
inline int& vlevel() { static int v = -1; return v; }
struct Sink {
  bool fatal;
  std::ostringstream os;
  explicit Sink(bool f, const char* file, int line, const char* what = nullptr) : fatal(f) {
    if (f) os << "[FATAL " << file << ":" << line << "] ";
    if (what) os << "Check failed: " << what << " ";
  }
  ~Sink() {
    std::string s = os.str();
    if (!s.empty() && s.back() != '\n') s.push_back('\n');
    std::cerr << s << std::flush;
    if (fatal) std::abort();
  }
  template <typename T> Sink& operator<<(const T& v) { os << v; return *this; }
  Sink& operator<<(std::ostream& (*f)(std::ostream&)) { f(os); return *this; }
};
struct Voidify { void operator&(const Sink&) {} };
template <typename T> T&& check_notnull(const char* file, int line, const char* expr, T&& t) {
  if (t == nullptr) { Sink(true, file, line, expr) << "must be non-NULL"; }
  return static_cast<T&&>(t);
}
}  // namespace ggnn_shim_log

namespace google {
inline void InitGoogleLogging(const char*) {}
inline void LogToStderr() {}
inline void InstallFailureSignalHandler() {}
inline void SetVLOGLevel(const char*, int level) { ggnn_shim_log::vlevel() = level; }
}  // namespace google

#define GGNN_SHIM_SEV_INFO false
#define GGNN_SHIM_SEV_WARNING false
#define GGNN_SHIM_SEV_ERROR false
#define GGNN_SHIM_SEV_FATAL true
#define GGNN_SHIM_SEV_DFATAL true

#define GGNN_SHIM_LOG_IF(cond, fatal, what) \
  !(cond) ? (void)0 : ::ggnn_shim_log::Voidify() & ::ggnn_shim_log::Sink(fatal, __FILE__, __LINE__, what)

#define LOG(sev) GGNN_SHIM_LOG_IF(true, GGNN_SHIM_SEV_##sev, nullptr)
#define LOG_IF(sev, cond) GGNN_SHIM_LOG_IF((cond), GGNN_SHIM_SEV_##sev, nullptr)
#define DLOG(sev) GGNN_SHIM_LOG_IF(false, GGNN_SHIM_SEV_##sev, nullptr)
#define DLOG_IF(sev, cond) GGNN_SHIM_LOG_IF(false && (cond), GGNN_SHIM_SEV_##sev, nullptr)
#define VLOG_IS_ON(n) (::ggnn_shim_log::vlevel() >= (n))
#define VLOG(n) GGNN_SHIM_LOG_IF(VLOG_IS_ON(n), false, nullptr)
#define VLOG_IF(n, cond) GGNN_SHIM_LOG_IF(VLOG_IS_ON(n) && (cond), false, nullptr)

#define CHECK(cond) GGNN_SHIM_LOG_IF(!(cond), true, #cond)
#define GGNN_SHIM_CHECK_OP(a, b, op) GGNN_SHIM_LOG_IF(!((a)op(b)), true, #a " " #op " " #b)
#define CHECK_EQ(a, b) GGNN_SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) GGNN_SHIM_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) GGNN_SHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) GGNN_SHIM_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) GGNN_SHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) GGNN_SHIM_CHECK_OP(a, b, >=)
#define CHECK_NOTNULL(p) ::ggnn_shim_log::check_notnull(__FILE__, __LINE__, #p, (p))
