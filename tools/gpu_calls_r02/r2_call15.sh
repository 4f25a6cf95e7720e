#!/usr/bin/env bash
# Round 2, call 15: ncu launch list of one bench step + ncu --set full captures of the kernels the roofline / DESIGN cite.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] ncu launch list of one bench step (4 decode tokens)"
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/r02_launches_final.csv \
  python bench.py --steps 1 --warmup 1 --no-parity --no-e2e --comparator none --new-tokens 4 > $O/c15_ncu_bench.log 2>&1
echo "rc=$?"; wc -l $O/r02_launches_final.csv
echo "== [2] ncu --set full: 4 consecutive ViT GEMM launches (roofline.traffic)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 8 --launch-count 4 -f -o $O/r02_gemm_full \
  python tools/phase_times.py > $O/c15_ncu_gemm.log 2>&1
echo "rc=$?"
echo "== [3] ncu --set full: decode-side kernels (weight streams, attention, epilogue, lm head)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"smallbatch_gemm_kernel|decode_attention_kernel|decode_residual" --launch-skip 96 --launch-count 8 -f -o $O/r02_decode_full \
  python tools/decode_only.py 3 > $O/c15_ncu_decode.log 2>&1
echo "rc=$?"
echo "== [4] ncu --set full: flash attention (final), patch embedding, quantised stream"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_tc_ --launch-skip 4 --launch-count 2 -f -o $O/r02_fa_final \
  python tools/attn_profile.py 0 > $O/c15_ncu_fa.log 2>&1
echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"patch_embed_kernel|smallbatch_gemm_quant_kernel|dequant_weights_kernel" --launch-count 6 -f -o $O/r02_misc_full \
  python tools/quant_profile.py > $O/c15_ncu_misc.log 2>&1
echo "rc=$?"; ls -la $O/*.ncu-rep
