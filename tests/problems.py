"""Problem set shared by the oracle and GPU parity tests.

Names and definitions follow the reference's benchmark problem list
(/root/reference/benchmark/benchmarks.jl:32-46) whose published values are the golden vectors
in tests/golden/reference_csv_fts.json.
"""
import math

from fts_oracle import FAMILIES as F

PI = math.pi

# function name in the reference CSV -> (dim, family id, params, low, up)
CSV_PROBLEMS = {
    "x^6 - 2x^3 + 0.5": (1, F["F1_POLY6"], None, -1.0, 1.0),
    "1/(1+25x^2)": (1, F["F1_RUNGE"], [25.0], -1.0, 1.0),
    "log(1-x)": (1, F["F1_LOG1MX"], None, -1.0, 1.0),
    "1/sqrt(1-x^2)": (1, F["F1_INVSQRT1MX2"], None, -1.0, 1.0),
    "sin^2(1000x)": (1, F["F1_SIN2"], [1000.0], -PI, PI),
    "exp(x+y)": (2, F["F2_EXP"], None, [-1.0, -1.0], [1.0, 1.0]),
    "x^2 + y^2": (2, F["F2_X2PY2"], None, [-1.0, -1.0], [1.0, 1.0]),
    "1/sqrt((1-x^2)(1-y^2))": (2, F["F2_INVSQRT"], None, [-1.0, -1.0], [1.0, 1.0]),
    "exp(x+y+z)": (3, F["F3_EXP"], None, [-1.0] * 3, [1.0] * 3),
    "x^2*y^2*z^2": (3, F["F3_X2Y2Z2"], None, [-1.0] * 3, [1.0] * 3),
    "1/sqrt((1-x^2)(1-y^2)(1-z^2))": (3, F["F3_INVSQRT"], None, [-1.0] * 3, [1.0] * 3),
}

# Relative tolerance of the oracle (libm/glibc transcendental functions) against the reference's
# published value (Julia's pure-Julia sinh/tanh/exp/sin): bit-exact or a few ulp except where
# the integrand amplifies 1-ulp node/function differences:
#   sin^2(1000x): d/dx amplifies a 1-ulp node difference by ~1000*pi   -> 3e-14
#   `_avx` rows of endpoint-singular integrands: @turbo contracts 1-x^2 into an FMA, which moves
#   the outermost terms by percents (reference avx vs scalar differ by 2.3e-14 themselves) -> 6e-14
CSV_RTOL = {
    ("sin^2(1000x)", "adaptive"): 3e-14, ("sin^2(1000x)", "quad"): 3e-14, ("sin^2(1000x)", "avx"): 3e-14,
    ("1/sqrt(1-x^2)", "avx"): 6e-14, ("1/sqrt((1-x^2)(1-y^2))", "avx"): 6e-14,
    ("1/sqrt((1-x^2)(1-y^2)(1-z^2))", "avx"): 6e-14,
    ("x^2*y^2*z^2", "adaptive"): 2e-15, ("x^2*y^2*z^2", "quad"): 2e-15,
}
CSV_RTOL_DEFAULT = 1e-15  # <= 4 ulp
CSV_RTOL_AVX_DEFAULT = 4e-15  # reassociated sums of the reference vs sequential/exact sums here
