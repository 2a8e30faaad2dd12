#include "fts_inst.cuh"
FTS_REGISTER_TPI(FTS_F3_LOG, F3Log, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_TPI(FTS_F3_INVSQRTABS, F3InvSqrtAbs, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_CTA(FTS_F3_LOG, F3Log, k_fixed3d_cta, k_adaptive3d_cta)
FTS_REGISTER_GRID(FTS_F3_LOG, F3Log, 3)
FTS_REGISTER_CTA(FTS_F3_INVSQRTABS, F3InvSqrtAbs, k_fixed3d_cta, k_adaptive3d_cta)
FTS_REGISTER_GRID(FTS_F3_INVSQRTABS, F3InvSqrtAbs, 3)
FTS_REGISTER_ORD(FTS_F3_LOG, F3Log, k_adaptive3d_ord)
FTS_REGISTER_ORD(FTS_F3_INVSQRTABS, F3InvSqrtAbs, k_adaptive3d_ord)
