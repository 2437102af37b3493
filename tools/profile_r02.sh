#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md commands; one GPU; every program below exits 0 without ncu first).
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of three eager Ours_Full steps (gpu__time_duration per launch; serialised, cold caches)
RAU_GRAPH=0 $NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/r02_launches_ours_full.csv python tools/phases.py ours_full > $O/r02_launches.log 2>&1
# 2. the encoder's persistent forward recurrence and the fused ATTLSTM cell (gate GEMMs): --set full
RAU_GRAPH=0 $NCU --set full --import-source on --kernel-name-base demangled -k regex:lstm_seq_kernel -c 2 -o $O/r02_lstm_seq python tools/phases.py ours_full > $O/r02_lstm_seq.log 2>&1
RAU_GRAPH=0 $NCU --set full --import-source on --kernel-name-base demangled -k 'regex:rows_gemm_kernel<\(int\)6' -c 2 -o $O/r02_epi_lstm python tools/phases.py ours_full > $O/r02_epi_lstm.log 2>&1
# 3. the encoder backward's per-step split-K dgrad
$NCU --set full --import-source on -k regex:rows_gemm_kernel -c 3 -o $O/r02_enc_dgrad python tools/enc_dgrad.py > $O/r02_enc_dgrad.log 2>&1
# 4. the HBM-bound kernels of an answering unit (pack, content logits, softmax + weighted sum, dp, dz): dram bytes + time
$NCU --set full -k 'regex:attn_rows|xprep_rows' -o $O/r02_hbm_kernels python tools/ncu_sweep.py mixed > $O/r02_hbm_kernels.log 2>&1
# 5. the i_embed product alone on 148 SMs in the default mode (the roofline kernel) was captured as tanh_mixed / here again after the 16-warp epilogue
$NCU --set full --import-source on -k regex:rows_gemm_kernel -c 2 -o $O/r02_tanh_mixed python tools/ncu_iembed.py mixed > $O/r02_tanh.log 2>&1
ls -la $O/*.ncu-rep
