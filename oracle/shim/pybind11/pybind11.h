// Stand-in for <pybind11/pybind11.h> — TEST INFRASTRUCTURE (oracle/_ref build only).  src/extend.cpp names
// pybind11 only to raise Python errors and to poll for Ctrl-C; outside an interpreter those are no-ops.
#pragma once
// <Python.h> (which the real header includes first) brings in the C headers below; with libstdc++ the C++
// <stdlib.h>/<math.h> wrappers put std::abs(float) into the global namespace, which src/extend.cpp:186 relies on
// (`abs(dist - prev_value)` on floats) — without them the call would bind to int abs(int).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdexcept>
#include <unordered_set>
namespace pybind11 {
struct key_error : std::runtime_error { key_error() : std::runtime_error("key_error") {} };
struct error_already_set : std::runtime_error { error_already_set() : std::runtime_error("error_already_set") {} };
}
inline int PyErr_CheckSignals() { return 0; }
