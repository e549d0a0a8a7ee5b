#!/bin/bash
# Builds the host-only translation units of libwvb200 (scene_host.cpp: OBJ reader + octree voxeliser;
# lrs_design.cpp: boundary filter design) with ASan + UBSan and runs tools/host_fuzz.py on them.
# No GPU needed. Output: profiles/r02_host_sanitizer.txt
set -e
cd "$(dirname "$0")/.."
out=gpurun_out/asan
mkdir -p $out
cat > $out/stub.cpp <<'EOC'
namespace wvb { void set_last_error(const char*, ...) {} }
EOC
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared -I include \
    -I wayverb_b200/csrc -I /usr/local/cuda/include -pthread wayverb_b200/csrc/scene_host.cpp \
    wayverb_b200/csrc/lrs_design.cpp $out/stub.cpp -o $out/libhost_asan.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
    ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 WVB_HOST_ASAN_LIB=$out/libhost_asan.so \
    python tools/host_fuzz.py 2>&1 | tee profiles/r02_host_sanitizer.txt
