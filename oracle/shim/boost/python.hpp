// TEST INFRASTRUCTURE ONLY -- build scaffolding for oracle/_ref.
//
// The reference's two native libraries (src/cpp/matchers/matchers.cpp,
// src/cpp/featextract/featextract.cpp) use Boost.Python for exactly three
// things: BOOST_PYTHON_MODULE(name), def("name", fn) and
// boost::python::numpy::initialize().  Boost is not installed in this image,
// so this header maps those three onto pybind11 and lets the *unmodified*
// reference sources compile where they lie under /root/reference.
// Nothing from the reference is copied; nothing here is product code.
#pragma once
#include <pybind11/pybind11.h>
#include <utility>

namespace boost {
namespace python {

namespace shim_detail {
inline pybind11::module_*& active_module() {
  static pybind11::module_* mod = nullptr;
  return mod;
}
// PyObject* parameters arrive as borrowed pybind11 handles; everything else
// (int, float) is passed through unchanged.
template <class T> struct param { using py_type = T; static T to_c(T v) { return v; } };
template <> struct param<PyObject*> {
  using py_type = pybind11::object;
  static PyObject* to_c(const pybind11::object& o) { return o.ptr(); }
};
}  // namespace shim_detail

// Functions that return a new ndarray reference (all the matchers / extractors).
template <class... Args>
void def(const char* name, PyObject* (*fn)(Args...)) {
  shim_detail::active_module()->def(
      name, [fn](typename shim_detail::param<Args>::py_type... args) {
        PyObject* out = fn(shim_detail::param<Args>::to_c(args)...);
        return pybind11::reinterpret_steal<pybind11::object>(out);
      });
}
// initthreads(): plain int return.
template <class... Args>
void def(const char* name, int (*fn)(Args...)) {
  shim_detail::active_module()->def(name, fn);
}

namespace numpy { inline void initialize() {} }

}  // namespace python
}  // namespace boost

#define BOOST_PYTHON_MODULE(modname)                                   \
  static void shim_module_body_##modname();                            \
  PYBIND11_MODULE(modname, shim_m) {                                   \
    boost::python::shim_detail::active_module() = &shim_m;             \
    shim_module_body_##modname();                                      \
  }                                                                    \
  static void shim_module_body_##modname()
