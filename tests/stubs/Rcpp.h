// tests/stubs/Rcpp.h -- NOT Rcpp.  A minimal stand-in with just enough of Rcpp's surface for
// `g++ -fsyntax-only integration/r_shim/oem_b200_shim.cpp` (tests/test_abi.py): it checks the shim's calls against the
// prototypes of include/oem_b200.h in an image that has neither R nor Rcpp.  Nothing here is ever linked or run.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

typedef struct SEXPREC *SEXP;
#define RcppExport extern "C"
#define BEGIN_RCPP try {
#define END_RCPP } catch (std::exception &e_) { (void)e_; } return (SEXP)0;
inline bool Rf_isS4(SEXP) { return false; }

namespace Rcpp {

struct exception : std::runtime_error { explicit exception(const char *m) : std::runtime_error(m) {} };
[[noreturn]] inline void stop(const char *m) { throw exception(m); }

template <typename T> struct Vec {
    std::vector<T> v;
    Vec() {}
    Vec(SEXP) {}
    explicit Vec(int n) : v((size_t)n) {}
    template <typename It> Vec(It a, It b) : v(a, b) {}
    T *begin() { return v.data(); }
    const T *begin() const { return v.data(); }
    long size() const { return (long)v.size(); }
    T &operator[](long i) { return v[(size_t)i]; }
    operator SEXP() const { return (SEXP)0; }
};
typedef Vec<double> NumericVector;
typedef Vec<int> IntegerVector;
typedef Vec<int> LogicalVector;
struct CharacterVector { CharacterVector() {} CharacterVector(SEXP) {} SEXP operator[](int) const { return (SEXP)0; } };

struct NumericMatrix {
    std::vector<double> v; int r = 0, c = 0;
    NumericMatrix(SEXP) {}
    NumericMatrix(int r_, int c_) : v((size_t)r_ * c_), r(r_), c(c_) {}
    int nrow() const { return r; }
    int ncol() const { return c; }
    double *begin() { return v.data(); }
    operator SEXP() const { return (SEXP)0; }
};

template <typename T> T as(SEXP) { return T(); }
template <typename T> T as(const CharacterVector &) { return T(); }

struct Proxy {
    template <typename T> Proxy &operator=(const T &) { return *this; }
    operator SEXP() const { return (SEXP)0; }
    template <typename T> operator Vec<T>() const { return Vec<T>(); }
};
struct NamedArg { template <typename T> NamedArg &operator=(const T &) { return *this; } };
inline NamedArg Named(const char *) { return NamedArg(); }

struct List {
    List() {}
    List(SEXP) {}
    explicit List(int) {}
    int size() const { return 0; }
    Proxy operator[](int) const { return Proxy(); }
    Proxy operator[](const char *) const { return Proxy(); }
    bool containsElementNamed(const char *) const { return false; }
    template <typename... A> static List create(const A &...) { return List(); }
    operator SEXP() const { return (SEXP)0; }
};
template <typename T> T as(const Proxy &) { return T(); }

struct S4 { S4(SEXP) {} Proxy slot(const char *) const { return Proxy(); } };
template <class T> struct PreserveStorage {};
template <typename T> void standard_delete_finalizer(T *obj) { delete obj; }
// same template signature as Rcpp's XPtr (storage policy, finalizer)
template <typename T, template <class> class StoragePolicy = PreserveStorage, void Finalizer(T *) = standard_delete_finalizer<T> >
struct XPtr {
    XPtr(SEXP) {}
    explicit XPtr(T *, bool = true) {}
    T *operator->() const { return (T *)0; }
    T *get() const { return (T *)0; }
    operator SEXP() const { return (SEXP)0; }
};

}  // namespace Rcpp
