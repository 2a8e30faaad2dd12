// CPU unit-test shim for the product's double-double arithmetic (csrc/fts_dd.cuh is
// __host__ __device__): exposes the operations on (hi, lo) pairs for tests/test_dd_arith.py.
#include "../fasttanhsinhquadrature.jl_b200/csrc/fts_dd.cuh"
using fts::dd;
extern "C" {
void t_dd_add(const double* a, const double* b, double* r) { dd v = dd(a[0], a[1]) + dd(b[0], b[1]); r[0] = v.hi; r[1] = v.lo; }
void t_dd_sub(const double* a, const double* b, double* r) { dd v = dd(a[0], a[1]) - dd(b[0], b[1]); r[0] = v.hi; r[1] = v.lo; }
void t_dd_mul(const double* a, const double* b, double* r) { dd v = dd(a[0], a[1]) * dd(b[0], b[1]); r[0] = v.hi; r[1] = v.lo; }
void t_dd_div(const double* a, const double* b, double* r) { dd v = dd(a[0], a[1]) / dd(b[0], b[1]); r[0] = v.hi; r[1] = v.lo; }
void t_dd_sqrt(const double* a, double* r) { dd v = fts::dd_sqrt(dd(a[0], a[1])); r[0] = v.hi; r[1] = v.lo; }
void t_dd_rsqrt(const double* a, double* r) { dd v = fts::dd_rsqrt(dd(a[0], a[1])); r[0] = v.hi; r[1] = v.lo; }
void t_dd_exp(const double* a, double* r) { dd v = fts::dd_exp(dd(a[0], a[1])); r[0] = v.hi; r[1] = v.lo; }
void t_dd_log(const double* a, double* r) { dd v = fts::dd_log(dd(a[0], a[1])); r[0] = v.hi; r[1] = v.lo; }
void t_dd_sin(const double* a, double* r) { dd v = fts::dd_sin(dd(a[0], a[1])); r[0] = v.hi; r[1] = v.lo; }
void t_dd_cos(const double* a, double* r) { dd v = fts::dd_cos(dd(a[0], a[1])); r[0] = v.hi; r[1] = v.lo; }
}
