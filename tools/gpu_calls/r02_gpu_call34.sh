#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/vae_profile.py > gpurun_out/r02ae_vae_profile.log 2>&1
cut -c1-230 gpurun_out/r02ae_vae_profile.log | head -60
