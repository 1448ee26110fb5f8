// Stand-in for boost::sort::parallel_stable_sort — TEST INFRASTRUCTURE (oracle/_ref build only).
// A stable sort has one defined result whatever the thread count: std::stable_sort gives the same permutation.
#pragma once
#include <algorithm>
#include <cstdint>
namespace boost { namespace sort {
template <class It, class Cmp> void parallel_stable_sort(It first, It last, Cmp cmp, uint32_t /*n_threads*/) {
    std::stable_sort(first, last, cmp);
}
}}
