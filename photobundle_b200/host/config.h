// config.h — minimal Boost-free ConfigFile with the reference's semantics (src/utils.h:282-379,
// src/utils.cc:143-212): "key = value" lines, '#' / '%' comments, case-insensitive keys,
// get<T>(name) throws when missing, get<T>(name, default) does not.
#ifndef PBA_HOST_CONFIG_H
#define PBA_HOST_CONFIG_H
#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace utils {
class ConfigFile {
 public:
  ConfigFile() {}
  explicit ConfigFile(const std::string& filename) {
    std::ifstream ifs(filename);
    if (!ifs.is_open()) throw std::runtime_error("ConfigFile: cannot open " + filename);
    parse(ifs);
  }
  void parse(std::istream& is) {
    std::string line;
    while (std::getline(is, line)) {
      const size_t c = line.find_first_of("#%");
      if (c != std::string::npos) line.erase(c);
      const size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      std::string k = strip(line.substr(0, eq)), v = strip(line.substr(eq + 1));
      if (!k.empty()) _data[lower(k)] = v;
    }
  }
  void set(const std::string& k, const std::string& v) { _data[lower(k)] = v; }
  template <class T> T get(const std::string& name) const {
    auto it = _data.find(lower(name));
    if (it == _data.end()) throw std::runtime_error("ConfigFile: no key " + name);
    return convert<T>(it->second);
  }
  template <class T> T get(const std::string& name, const T& def) const {
    auto it = _data.find(lower(name));
    if (it == _data.end()) return def;
    try { return convert<T>(it->second); } catch (...) { return def; }
  }
 private:
  std::map<std::string, std::string> _data;
  static std::string lower(std::string s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); }); return s; }
  static std::string strip(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
  }
  template <class T> static T convert(const std::string& v) {
    std::istringstream ss(v);
    T out;
    if (!(ss >> out)) throw std::runtime_error("ConfigFile: bad value " + v);
    return out;
  }
};
template <> inline std::string ConfigFile::convert<std::string>(const std::string& v) { return v; }
}  // namespace utils
#endif
