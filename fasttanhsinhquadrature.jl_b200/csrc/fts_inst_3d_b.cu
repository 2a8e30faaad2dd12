#include "fts_inst.cuh"
FTS_REGISTER_TPI(FTS_F3_INVSQRT, F3InvSqrt, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_TPI(FTS_F3_RAT, F3Rat, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_CTA(FTS_F3_INVSQRT, F3InvSqrt, k_fixed3d_cta, k_adaptive3d_cta)
FTS_REGISTER_GRID(FTS_F3_INVSQRT, F3InvSqrt, 3)
FTS_REGISTER_CTA(FTS_F3_RAT, F3Rat, k_fixed3d_cta, k_adaptive3d_cta)
FTS_REGISTER_GRID(FTS_F3_RAT, F3Rat, 3)
FTS_REGISTER_ORD(FTS_F3_INVSQRT, F3InvSqrt, k_adaptive3d_ord)
FTS_REGISTER_ORD(FTS_F3_RAT, F3Rat, k_adaptive3d_ord)
