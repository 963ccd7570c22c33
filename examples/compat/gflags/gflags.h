// Minimal stand-in for <gflags/gflags.h> (gflags is not installed in this image), just enough to compile the
// REFERENCE's unmodified examples/cpp-and-cuda/ggnn_benchmark.cpp against this repo's headers:
// DEFINE_{string,uint32,double,bool}, --name=value / --name value / --[no]name parsing, usage + version strings.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <string>

namespace gflags {
namespace detail {
struct Flag {
  std::string help, dflt;
  bool is_bool;
  std::function<bool(const std::string&)> set;
};
inline std::map<std::string, Flag>& registry()
{
  static std::map<std::string, Flag> r;
  return r;
}
inline std::string& usage()
{
  static std::string u;
  return u;
}
struct Registrar {
  Registrar(const char* name, const char* help, std::string dflt, bool is_bool, std::function<bool(const std::string&)> set)
  {
    registry()[name] = Flag{help, std::move(dflt), is_bool, std::move(set)};
  }
};
}  // namespace detail

inline void SetUsageMessage(const std::string& u) { detail::usage() = u; }
inline void SetVersionString(const std::string&) {}
inline void ShutDownCommandLineFlags() {}

inline uint32_t ParseCommandLineFlags(int* argc, char*** argv, bool /*remove_flags*/)
{
  auto fail = [](const std::string& msg) {
    std::cerr << "ERROR: " << msg << "\n";
    std::exit(1);
  };
  for (int i = 1; i < *argc; ++i) {
    std::string a = (*argv)[i];
    if (a == "--help" || a == "-help" || a == "-h") {
      std::cout << detail::usage() << "\n\nFlags:\n";
      for (const auto& [name, f] : detail::registry()) std::cout << "  --" << name << " (" << f.help << ") default: " << f.dflt << "\n";
      std::exit(0);
    }
    if (a.rfind("--", 0) == 0) a = a.substr(2);
    else if (a.rfind("-", 0) == 0) a = a.substr(1);
    else fail("unexpected argument '" + a + "'");
    std::string value;
    bool has_value = false;
    if (const size_t eq = a.find('='); eq != std::string::npos) {
      value = a.substr(eq + 1);
      a = a.substr(0, eq);
      has_value = true;
    }
    auto& reg = detail::registry();
    auto it = reg.find(a);
    if (it == reg.end() && a.rfind("no", 0) == 0 && reg.count(a.substr(2)) && reg[a.substr(2)].is_bool && !has_value) {
      it = reg.find(a.substr(2));
      value = "false";
      has_value = true;
    }
    if (it == reg.end()) fail("unknown command line flag '" + a + "'");
    if (!has_value) {
      if (it->second.is_bool) value = "true";
      else if (i + 1 < *argc) value = (*argv)[++i];
      else fail("flag '--" + a + "' is missing its argument");
    }
    if (!it->second.set(value)) fail("illegal value '" + value + "' specified for flag '" + a + "'");
  }
  return 1;
}
}  // namespace gflags

#define GGNN_COMPAT_DEFINE_FLAG(type, name, dflt, help, is_bool, parse)                                   \
  type FLAGS_##name = dflt;                                                                                \
  static ::gflags::detail::Registrar gflags_registrar_##name(#name, help, #dflt, is_bool,                 \
                                                             [](const std::string& s) -> bool { parse; })

#define DEFINE_string(name, dflt, help) GGNN_COMPAT_DEFINE_FLAG(std::string, name, dflt, help, false, FLAGS_##name = s; return true)
#define DEFINE_uint32(name, dflt, help)                                                                   \
  GGNN_COMPAT_DEFINE_FLAG(uint32_t, name, dflt, help, false, char* end = nullptr; const unsigned long v = std::strtoul(s.c_str(), &end, 10); \
                          if (s.empty() || *end) return false; FLAGS_##name = static_cast<uint32_t>(v); return true)
#define DEFINE_double(name, dflt, help)                                                                   \
  GGNN_COMPAT_DEFINE_FLAG(double, name, dflt, help, false, char* end = nullptr; const double v = std::strtod(s.c_str(), &end);   \
                          if (s.empty() || *end) return false; FLAGS_##name = v; return true)
#define DEFINE_bool(name, dflt, help)                                                                     \
  GGNN_COMPAT_DEFINE_FLAG(bool, name, dflt, help, true, if (s == "true" || s == "1" || s == "yes") FLAGS_##name = true;          \
                          else if (s == "false" || s == "0" || s == "no") FLAGS_##name = false; else return false; return true)
