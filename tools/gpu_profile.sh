#!/bin/bash
# One gpurun call: per-kernel micro-benchmark, ncu launch list of one UNet evaluation + decode, and
# `ncu --set full` captures of the three hot kernels (each only after its plain run exited 0).
set -u
TAG=${1:-r02}
K='regex:^(gemm_|attn2?_tc|gn_|layernorm|softmax_rows|fill_zero|nchw_f32|nhwc_to|upsample2x|conv_direct|small_linear|cfg_ddpm|vae_|f32_to_bf16|axpby|image_to|uint8_to|clip_embed|matmul_f64|resample_u8)'
python tools/kernel_bench.py --graph --attn-mode 3 --json gpurun_out/${TAG}_kernel_bench.json > gpurun_out/${TAG}_kernel_bench.log 2>&1
python bench.py --profile-only > gpurun_out/${TAG}_po.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile-only > gpurun_out/${TAG}_ncu_launches.log 2>&1
for c in conv3x3_res_320_320_64 conv3x3_mergedskip640_320_320_64 linear_proj_65536x320x320 linear_tok_65536x320x320 attn_self_S4096; do
  python tools/kernel_bench.py --attn-mode 3 --only $c --iters 1 --warmup 1 > gpurun_out/${TAG}_plain_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k 'regex:^(gemm_tc|attn2?_tc)' -s 1 -c 1 -f \
      -o gpurun_out/${TAG}_prof_$c python tools/kernel_bench.py --attn-mode 3 --only $c --iters 1 --warmup 1 \
      > gpurun_out/${TAG}_ncu_$c.log 2>&1
  python tools/ncu_src.py gpurun_out/${TAG}_prof_$c.ncu-rep 8 > gpurun_out/${TAG}_ncusum_$c.txt 2>&1
done
# HBM-bound side: GroupNorm apply (+SiLU) and LayerNorm, 2nd matching launch of the named kernel
for pair in "groupnorm_x_320_64:gn_apply_kernel" "groupnorm_hidparts_320_64:gn_apply_kernel" "layernorm_65536x320:layernorm_f32_kernel" "layernorm_h16_65536x320:layernorm_h16_kernel"; do
  c=${pair%%:*}; k=${pair##*:}
  python tools/kernel_bench.py --only $c --iters 1 --warmup 1 > gpurun_out/${TAG}_plain_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:^(void )?$k" -s 1 -c 1 -f \
      -o gpurun_out/${TAG}_prof_$c python tools/kernel_bench.py --only $c --iters 1 --warmup 1 \
      > gpurun_out/${TAG}_ncu_$c.log 2>&1
  python tools/ncu_src.py gpurun_out/${TAG}_prof_$c.ncu-rep 8 > gpurun_out/${TAG}_ncusum_$c.txt 2>&1
done
# gpurun brings back at most 64 MiB: the text summaries above travel, of the ~10 MB reports only the three hot ones
for c in conv3x3_res_320_320_64 linear_proj_65536x320x320 groupnorm_x_320_64 groupnorm_hidparts_320_64 layernorm_65536x320 layernorm_h16_65536x320; do
  rm -f gpurun_out/${TAG}_prof_$c.ncu-rep
done
tail -3 gpurun_out/${TAG}_kernel_bench.log
ls -la gpurun_out | tail -20
