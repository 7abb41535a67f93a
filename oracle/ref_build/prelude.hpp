// Force-included (-include) ahead of every reference TU: the reference relies on transitive
// STL includes that GCC 13 no longer provides (include/constants.hpp:96 needs <array>,
// include/aggregats/aggregat.hpp:64 needs <unordered_map>, ...).  Test infrastructure only.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
