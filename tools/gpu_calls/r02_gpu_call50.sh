#!/bin/bash
# round 2, call 50: ncu launch list of the LAST commit (after the 16x16 GroupNorm change)
set -u
mkdir -p gpurun_out
TAG=r02av
K='regex:^(gemm_|attn2?_tc|gn_|layernorm|softmax_rows|fill_zero|nchw_f32|nhwc_to|upsample2x|conv_direct|small_linear|cfg_ddpm|vae_|f32_to_bf16|axpby|image_to|uint8_to|clip_embed|matmul_f64|resample_u8|copy_bytes)'
timeout 300 python bench.py --profile-only > gpurun_out/${TAG}_po.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile-only > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv | head -30
