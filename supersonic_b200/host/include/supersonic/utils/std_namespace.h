// Forwarding header for the reference's supersonic/utils/std_namespace.h (a list of using-declarations for the
// std names the reference's sources use unqualified); client code includes it next to the standard headers.
#ifndef SUPERSONIC_B200_HOST_UTILS_STD_NAMESPACE_H_
#define SUPERSONIC_B200_HOST_UTILS_STD_NAMESPACE_H_
#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>
using std::copy;
using std::make_pair;
using std::map;
using std::max;
using std::min;
using std::multimap;
using std::pair;
using std::set;
using std::sort;
using std::string;
using std::swap;
using std::vector;
#endif
