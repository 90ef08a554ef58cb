#!/bin/bash
# `ncu --set full` captures of the three benched kernels at the current sources (run under gpurun); digests: profiles/ncu_digest.py
cd $GRAFT_REPO_ROOT
bash scripts/gpu_ncu.sh AD_64800_R12_GF256 592; mv gpurun_out/prof_AD_64800_R12_GF256.ncu-rep gpurun_out/prof_q256_bubble.ncu-rep
bash scripts/gpu_ncu.sh AD_64800_R12_GF256 296 --ecn syndrome; mv gpurun_out/prof_AD_64800_R12_GF256.ncu-rep gpurun_out/prof_q256_syndrome.ncu-rep
bash scripts/gpu_ncu.sh Ahmed_64800_R34_GF16 1184; mv gpurun_out/prof_Ahmed_64800_R34_GF16.ncu-rep gpurun_out/prof_q16_bubble.ncu-rep
