#!/bin/bash
mkdir -p gpurun_out
python tools/debug_ooc.py 21 20 6 2 0 1 0 > gpurun_out/dbg_fresh.txt 2>&1; tail -4 gpurun_out/dbg_fresh.txt
python tools/debug_ooc.py 21 20 6 2 0 1 4096 > gpurun_out/dbg_dirty.txt 2>&1; grep -c differs gpurun_out/dbg_dirty.txt; grep differs gpurun_out/dbg_dirty.txt | head -5; tail -2 gpurun_out/dbg_dirty.txt
compute-sanitizer --tool memcheck python tools/debug_ooc.py 12 12 5 2 0 1 64 0 > gpurun_out/san_mem.txt 2>&1; tail -8 gpurun_out/san_mem.txt
compute-sanitizer --tool initcheck python tools/debug_ooc.py 12 12 5 2 0 1 64 0 > gpurun_out/san_init.txt 2>&1; grep -A12 "Uninitialized" gpurun_out/san_init.txt | head -60; tail -4 gpurun_out/san_init.txt
