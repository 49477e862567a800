#!/bin/bash
# round 2, GPU call 18 (last minutes of the budget): the NHWC transform's new default (128 pixels per CTA) under the VAE / conv parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
timeout 90 python -m pytest tests/test_gpu_parity_config.py tests/test_gpu_models.py tests/test_gpu_ops.py -q -m gpu -x -k "vae_decode_512 or vae_decoder_vs or resblock or upsample_conv or test_conv_2d" 2>&1 | tail -4 | tee gpurun_out/r2c18_tests.log
