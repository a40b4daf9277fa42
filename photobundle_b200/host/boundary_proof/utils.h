// declaration-only stand-in for utils::ProgramOptions (src/utils.h:381-470); utils::ConfigFile is the
// host side's own (../config.h), the one PhotometricBundleAdjustment::Options is constructed from
#ifndef PBA_PROOF_UTILS_H
#define PBA_PROOF_UTILS_H
#include <string>
#include "config.h"
#include "debug.h"
namespace utils {
class ProgramOptions {
 public:
  ProgramOptions(std::string name = "ProgramOptions");
  ProgramOptions& operator()(const char* key, const char* msg);
  template <class T> ProgramOptions& operator()(const char* key, T default_value, const char* msg);
  ProgramOptions& parse(int argc, char** argv);
  template <class T> T get(std::string key) const;
  bool hasOption(std::string) const;
};
}  // namespace utils
#endif
