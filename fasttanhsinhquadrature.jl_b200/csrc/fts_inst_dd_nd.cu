// Double64 (double-double) 2D / 3D kernel set: thread per integral (the reference's adaptive_integrate_2D/3D and
// integrate2D/3D are generic in T<:Real and are exercised with Double64; quad.jl:117,155).
#include "fts_inst.cuh"

#define FTS_REGISTER_DD_ND(FAM, FUNCTOR, FIXED_K, ADAPT_K)                                                    \
    static const fts_host::KernelEntry k_ddn_entries_##FUNCTOR[] = {                                         \
        {FAM, FTS_DD64, fts_host::KID_FIXED_TPI, 0, FTS_KPTR(fts::FIXED_K<fts::dd, fts::FUNCTOR, false>)},   \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::ADAPT_K<fts::dd, fts::FUNCTOR, false>)},   \
    };                                                                                                        \
    static fts_host::Registrar k_ddn_reg_##FUNCTOR(k_ddn_entries_##FUNCTOR, 2);

#define FTS_REGISTER_DD_ND_ADAPT(FAM, FUNCTOR, ADAPT_K)                                                       \
    static const fts_host::KernelEntry k_ddn_entries_##FUNCTOR[] = {                                         \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::ADAPT_K<fts::dd, fts::FUNCTOR, false>)},   \
    };                                                                                                        \
    static fts_host::Registrar k_ddn_reg_##FUNCTOR(k_ddn_entries_##FUNCTOR, 1);

FTS_REGISTER_DD_ND(FTS_F2_POLY2, F2Poly2, k_fixed2d_tpi, k_adaptive2d_tpi)
FTS_REGISTER_DD_ND(FTS_F2_EXP, F2Exp, k_fixed2d_tpi, k_adaptive2d_tpi)
FTS_REGISTER_DD_ND(FTS_F2_X2PY2, F2X2pY2, k_fixed2d_tpi, k_adaptive2d_tpi)
FTS_REGISTER_DD_ND(FTS_F2_INVSQRT, F2InvSqrt, k_fixed2d_tpi, k_adaptive2d_tpi)
FTS_REGISTER_DD_ND(FTS_F2_INVSQRTABS, F2InvSqrtAbs, k_fixed2d_tpi, k_adaptive2d_tpi)
FTS_REGISTER_DD_ND_ADAPT(FTS_C2_INVSQRT_BMX_YMC, C2InvSqrtBmxYmc, k_adaptive2d_tpi)
FTS_REGISTER_DD_ND_ADAPT(FTS_C2_INVSQRT_ALL, C2InvSqrtAll, k_adaptive2d_tpi)
FTS_REGISTER_DD_ND(FTS_F3_POLY2, F3Poly2, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_DD_ND(FTS_F3_EXP, F3Exp, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_DD_ND(FTS_F3_X2Y2Z2, F3X2Y2Z2, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_DD_ND(FTS_F3_INVSQRT, F3InvSqrt, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_DD_ND(FTS_F3_RAT, F3Rat, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_DD_ND(FTS_F3_LOG, F3Log, k_fixed3d_tpi, k_adaptive3d_tpi)
FTS_REGISTER_DD_ND(FTS_F3_INVSQRTABS, F3InvSqrtAbs, k_fixed3d_tpi, k_adaptive3d_tpi)
