#!/bin/bash
# SASS evidence of the bulk-async (TMA) staging, mbarriers and programmatic dependent launch in the 2D cooperative
# kernels: counts of the mnemonics per kernel, from the objects of the current build.  usage: tools/sass_evidence.sh > profiles/r2_tma_sass.md
cd "$(dirname "$0")/../fasttanhsinhquadrature.jl_b200"
echo "# SASS evidence: TMA (cp.async.bulk) staging + mbarrier + programmatic dependent launch in the 2D cooperative kernels"
echo
echo 'Source: `cuobjdump -sass -fun <kernel> build/fts_inst_cmpl.o` of the committed sources (`tools/sass_evidence.sh`), sm_100a.'
echo '`UBLKCP.S.G` = `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`; `SYNCS.ARRIVE.TRANS64` = `mbarrier.arrive.expect_tx`;'
echo '`SYNCS.PHASECHK.TRANS64.TRYWAIT` = `mbarrier.try_wait.parity`; `ELECT` picks the issuing lane; `PREEXIT` = `griddepcontrol.launch_dependents`, `ACQBULK` = `griddepcontrol.wait` (programmatic dependent launch); `LDS.128` = the vector'
echo 'loads of the x-side records in the warp kernel.'
for k in k_adaptive2d_warp k_adaptive2d_cta; do
  for fam in 15C2InvSqrtBmxYmc; do
    fn=$(cuobjdump -elf build/fts_inst_cmpl.o 2>/dev/null | grep -o "_ZN3fts[0-9]*${k}IdNS_${fam}EEEvNS_4ArgsIT_EE" | head -1)
    echo
    echo "## \`$k<double, C2InvSqrtBmxYmc>\`"
    echo
    echo '| count | mnemonic |'
    echo '|---|---|'
    cuobjdump -sass -fun "$fn" build/fts_inst_cmpl.o | grep -oE "UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|ACQBULK[.A-Z0-9]*|ELECT|LDS\.128|ACQBLK|PREEXIT|MUFU\.RSQ64H" | sort | uniq -c | sort -rn | awk '{print "| " $1 " | `" $2 "` |"}'
    echo
    echo "registers / stack: $(grep -h -A3 "$fn" build/fts_inst_cmpl.ptxas.log | grep -E "Used|stack frame" | tr '\n' ' ' | sed 's/ptxas info *: //g')"
  done
done
