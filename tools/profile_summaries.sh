#!/bin/bash
# Developer tool: turn the captures of a GPU call (tools/gpu/r2_call47.sh: gpurun_out/<tag>_counts.csv, _counts_full.csv,
# _prof_fwd.ncu-rep, _prof_bwd.ncu-rep) into the tracked summaries under profiles/ (kernel_counts.json, the operand model,
# the forward / backward summaries).  usage: tools/profile_summaries.sh r2c47 47     (run from the repository root, CPU only)
set -e
TAG=${1:?tag}; CALL=${2:?call number}
O=gpurun_out; W=$(mktemp -d); LIB=${STE_PROFILE_LIB:-ship_track_estimators_b200/csrc/libste_ukf.so}   # the library the captures were taken with
TRACKS=75776; STEPS=64; WS=$((TRACKS / 32 * STEPS))
( cd $W && cuobjdump -xelf all $OLDPWD/$LIB > /dev/null && nvdisasm -gi -c *.sm_100a.cubin > disasm_gi.txt )
raw() { ncu -i $O/${TAG}_prof_$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_line.py; }
for k in fwd bwd; do ncu -i $O/${TAG}_prof_$k.ncu-rep --page source --csv --print-source sass > $W/sass_$k.csv 2>/dev/null; done
{
  python tools/fp64_operand_model.py $W/sass_fwd.csv ukf_forward_kernel $WS $LIB ukf_forward_kernelILb1ELb0
  python tools/fp64_operand_model.py $W/sass_bwd.csv urtss_backward_kernel $WS $LIB urtss_backward_kernelILb1
  sed -n '/^DFMA reg operands/,$p' profiles/r02_fp64_operand_model.txt      # the issue-interval probes (GPU call 27) are kept
} > $W/operand_model.txt
cp $W/operand_model.txt profiles/r02_fp64_operand_model.txt
RF_F=$(grep -m1 "pipe + register-file cycles" $W/operand_model.txt | awk '{print $NF}')
python tools/ncu_counts.py $O/${TAG}_counts.csv $TRACKS $STEPS --full-cov $O/${TAG}_counts_full.csv \
    --note "GPU call $CALL (round's last tree)" > $W/kernel_counts.json
for k in fwd bwd; do
  if [ $k = fwd ]; then NAME=forward; SUB="ukf_forward_kernelILb1ELb0|ukf_forward_kernel"; else NAME=backward; SUB="urtss_backward_kernelILb1|urtss_backward_kernel"; fi
  {
    echo "Round 2, $NAME kernel of the round's last tree (GPU call $CALL): ncu --set full --clock-control none --import-source on, tools/quick_perf.py --tracks $TRACKS --steps $STEPS --packed"
    echo "($((TRACKS / 32)) warps x $STEPS steps = $WS warp-steps per launch), one B200.  Numbers are warp-level instructions per warp-step (= per track-step per thread)."
    raw $k
    echo "FP64 operand model (tools/fp64_operand_model.py, operands from cuobjdump): profiles/r02_fp64_operand_model.txt."
    echo
    python tools/ncu_sass_mix.py $W/sass_$k.csv $WS
    echo
    python tools/ncu_attrib.py $W/sass_$k.csv $W/disasm_gi.txt "$SUB" $WS
  } > profiles/r02_${NAME}_summary.txt
done
echo "$W/kernel_counts.json (review, then copy to profiles/kernel_counts.json); operand-model forward cycles: $RF_F"
