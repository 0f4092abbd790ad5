#!/bin/bash
cd /root/repo
timeout 300 python scratch/rs_bench.py 2>&1 | tail -20
