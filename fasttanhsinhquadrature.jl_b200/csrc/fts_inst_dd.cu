// Double64 (double-double) kernel set: thread-per-integral, reference order (BASELINE config E).
#include "fts_inst.cuh"

#define FTS_REGISTER_DD_1D(FAM, FUNCTOR)                                                                     \
    static const fts_host::KernelEntry k_dd_entries_##FUNCTOR[] = {                                         \
        {FAM, FTS_DD64, fts_host::KID_FIXED_TPI, 0, FTS_KPTR(fts::k_fixed1d_tpi<fts::dd, fts::FUNCTOR, false>)},   \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::k_adaptive1d_tpi<fts::dd, fts::FUNCTOR, false>)}, \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_TAIL, 0, FTS_KPTR(fts::k_adaptive1d_tail<fts::dd, fts::FUNCTOR, false, false>)}, \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_GRP, 0, FTS_KPTR(fts::k_adaptive1d_grp<fts::dd, fts::FUNCTOR, false, 4>)}, \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_GRP, 1, FTS_KPTR(fts::k_adaptive1d_grp<fts::dd, fts::FUNCTOR, false, 2>)}, \
    };                                                                                                       \
    static fts_host::Registrar k_dd_reg_##FUNCTOR(k_dd_entries_##FUNCTOR, 5);

#define FTS_REGISTER_DD_C1(FAM, FUNCTOR)                                                                     \
    static const fts_host::KernelEntry k_dd_entries_##FUNCTOR[] = {                                         \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::k_adaptive1d_cmpl_tpi<fts::dd, fts::FUNCTOR, false>)}, \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_TAIL, 0, FTS_KPTR(fts::k_adaptive1d_tail<fts::dd, fts::FUNCTOR, false, true>)}, \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_GRP, 0, FTS_KPTR(fts::k_adaptive1d_grp<fts::dd, fts::FUNCTOR, true, 4>)}, \
        {FAM, FTS_DD64, fts_host::KID_ADAPT_GRP, 1, FTS_KPTR(fts::k_adaptive1d_grp<fts::dd, fts::FUNCTOR, true, 2>)}, \
    };                                                                                                       \
    static fts_host::Registrar k_dd_reg_##FUNCTOR(k_dd_entries_##FUNCTOR, 4);

FTS_REGISTER_DD_1D(FTS_F1_POLY7, F1Poly7)
FTS_REGISTER_DD_1D(FTS_F1_EXP, F1Exp)
FTS_REGISTER_DD_1D(FTS_F1_GAUSS, F1Gauss)
FTS_REGISTER_DD_1D(FTS_F1_RUNGE, F1Runge)
FTS_REGISTER_DD_1D(FTS_F1_LOG1MX, F1Log1mx)
FTS_REGISTER_DD_1D(FTS_F1_LOG1PX, F1Log1px)
FTS_REGISTER_DD_1D(FTS_F1_INVSQRT1MX2, F1InvSqrt1mx2)
FTS_REGISTER_DD_1D(FTS_F1_POLY6, F1Poly6)
FTS_REGISTER_DD_1D(FTS_F1_INVSQRTABS, F1InvSqrtAbs)
FTS_REGISTER_DD_C1(FTS_C1_INVSQRT, C1InvSqrt)
FTS_REGISTER_DD_C1(FTS_C1_LOG_BMX, C1LogBmx)
FTS_REGISTER_DD_C1(FTS_C1_LOG_XMA, C1LogXma)
FTS_REGISTER_DD_C1(FTS_C1_EXP_INV_BMX, C1ExpInvBmx)
FTS_REGISTER_DD_C1(FTS_C1_EXP_PLUS, C1ExpPlus)
