#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_train_demo.jsonl
for m in bf16 tf32; do
  for lr in 5e-6 5e-5; do
    timeout 600 python tools/train_demo.py --mode $m --epochs 10 --lr $lr >> gpurun_out/r2_train_demo.jsonl 2> gpurun_out/r2_train_demo.err
    echo "train_demo $m lr=$lr rc=$?"
  done
done
cut -c1-420 gpurun_out/r2_train_demo.jsonl
