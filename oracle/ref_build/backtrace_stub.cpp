// Replaces src/exceptions.cpp (libunwind v1.5 is absent, reference CMakeLists.txt:70-80).
// Test infrastructure only.
#include <string>
std::string Backtrace(int) { return {}; }
