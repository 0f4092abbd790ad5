#!/bin/bash
# compile scratch/one.cu -> cubin, dump SASS to scratch/one.sass, print summary
cd /root/repo/scratch
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 186 -cubin -Xptxas -v -o one.cubin one.cu 2>&1 | grep -E "registers|spill|error"
cuobjdump -sass one.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\*[0-9a-f]\{16\}\*/##' > one.sass
wc -l one.sass
