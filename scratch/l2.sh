#!/bin/bash
cd "$(dirname "$0")/.."
scratch/launches.sh cfg3 | tail -14
scratch/launches.sh cfg4 --ranks phylum,genus,species --mode above --samples 8 | tail -12
scratch/launches.sh cfg2 | tail -8
