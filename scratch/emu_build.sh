#!/bin/bash
# build the host emulation of the device math (tests/hostemu) into $1 (default /tmp/libemu.so), extra flags after
OUT=${1:-/tmp/libemu.so}; shift
R=/root/repo
g++ -O2 -fopenmp -shared -fPIC -std=c++17 -DPISAB_HOST_EMU -include $R/tests/hostemu/cuda_shim.h -I$R/pisa_b200/csrc -I$R/include "$@" -o $OUT $R/tests/hostemu/emu.cpp
