// Minimal stand-in for <glog/logging.h> (glog is not installed in this image), just enough to compile the
// REFERENCE's unmodified example programs against this repo's headers: LOG / VLOG / CHECK* stream macros.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace google {
namespace compat {
inline int& vlog_level()
{
  static int level = 0;
  return level;
}
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
// collects one message; prints it (and aborts for FATAL) when it goes out of scope
class Message {
 public:
  Message(Severity s, const char* file, int line, bool enabled = true) : severity_{s}, enabled_{enabled}
  {
    static constexpr char tags[] = {'I', 'W', 'E', 'F'};
    if (enabled_) stream_ << tags[s] << ' ' << file << ':' << line << "] ";
  }
  [[noreturn]] void die()
  {
    std::cerr << stream_.str() << std::endl;
    std::abort();
  }
  ~Message()
  {
    if (severity_ == FATAL) die();
    if (enabled_) std::cerr << stream_.str() << std::endl;
  }
  std::ostream& stream() { return stream_; }

 private:
  Severity severity_;
  bool enabled_;
  std::ostringstream stream_;
};
// FATAL messages never return: lets `LOG(FATAL) << ...;` end a non-void lambda like glog's does
class FatalMessage : public Message {
 public:
  FatalMessage(const char* file, int line) : Message(FATAL, file, line) {}
  [[noreturn]] ~FatalMessage() { die(); }
};
struct Voidify {
  void operator&(std::ostream&) {}
};
}  // namespace compat
inline void InitGoogleLogging(const char*) {}
inline void LogToStderr() {}
inline void InstallFailureSignalHandler() {}
inline void SetVLOGLevel(const char*, int level) { compat::vlog_level() = level; }
}  // namespace google

#define GGNN_COMPAT_LOG_INFO ::google::compat::Message(::google::compat::INFO, __FILE__, __LINE__)
#define GGNN_COMPAT_LOG_WARNING ::google::compat::Message(::google::compat::WARNING, __FILE__, __LINE__)
#define GGNN_COMPAT_LOG_ERROR ::google::compat::Message(::google::compat::ERROR, __FILE__, __LINE__)
#define GGNN_COMPAT_LOG_FATAL ::google::compat::FatalMessage(__FILE__, __LINE__)
#define LOG(severity) GGNN_COMPAT_LOG_##severity.stream()
#define VLOG_IS_ON(n) ((n) <= ::google::compat::vlog_level())
#define VLOG(n) !VLOG_IS_ON(n) ? (void)0 : ::google::compat::Voidify() & LOG(INFO)
#define CHECK(cond) (cond) ? (void)0 : ::google::compat::Voidify() & LOG(FATAL) << "Check failed: " #cond " "
#define GGNN_COMPAT_CHECK_OP(a, b, op) CHECK((a)op(b))
#define CHECK_EQ(a, b) GGNN_COMPAT_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) GGNN_COMPAT_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) GGNN_COMPAT_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) GGNN_COMPAT_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) GGNN_COMPAT_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) GGNN_COMPAT_CHECK_OP(a, b, >=)
