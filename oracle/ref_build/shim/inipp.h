// Stand-in for the un-vendored `inipp` header (reference CMakeLists.txt:26-35 pulls
// https://gitlab.coria-cfd.fr/MCAC/inipp.git tag no_tests, a fork of mcmtroffaes/inipp).
// TEST INFRASTRUCTURE ONLY: lets the unmodified reference sources under /root/reference
// compile in this image.  Written from the call sites in
// src/physical_model/physical_model.cpp:27,101-187,271-272; no arithmetic lives here.
#pragma once
#include <istream>
#include <map>
#include <ostream>
#include <sstream>
#include <string>

namespace inipp {
template <typename CharT>
class Ini {
  public:
    using String = std::basic_string<CharT>;
    using Section = std::map<String, String>;
    std::map<String, Section> sections;

    static String trim(const String &s) {
        size_t b = s.find_first_not_of(" \t\r\n");
        if (b == String::npos) return String();
        size_t e = s.find_last_not_of(" \t\r\n");
        return s.substr(b, e - b + 1);
    }
    void parse(std::basic_istream<CharT> &is) {
        String line, section;
        while (std::getline(is, line)) {
            line = trim(line);
            if (line.empty() || line[0] == ';' || line[0] == '#') continue;
            if (line[0] == '[') {
                size_t end = line.find(']');
                if (end != String::npos) section = trim(line.substr(1, end - 1));
                continue;
            }
            size_t eq = line.find('=');
            if (eq == String::npos) continue;
            String key = trim(line.substr(0, eq));
            String val = trim(line.substr(eq + 1));
            if (sections[section].count(key) == 0) sections[section][key] = val;
        }
    }
    void interpolate() {}
    void generate(std::basic_ostream<CharT> &os) const {
        for (const auto &sec : sections) {
            os << "[" << sec.first << "]" << std::endl;
            for (const auto &kv : sec.second) os << kv.first << "=" << kv.second << std::endl;
            os << std::endl;
        }
    }
};
template <typename CharT, typename T>
inline bool extract(const std::basic_string<CharT> &value, T &dst) {
    CharT c;
    std::basic_istringstream<CharT> is{value};
    T result;
    if ((is >> std::boolalpha >> result) && !(is >> c)) {
        dst = result;
        return true;
    }
    return false;
}
template <typename CharT>
inline bool extract(const std::basic_string<CharT> &value, std::basic_string<CharT> &dst) {
    dst = value;
    return true;
}
}  // namespace inipp
