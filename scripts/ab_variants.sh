#!/usr/bin/env bash
# A/B of the compile-time kernel experiments against the default library.
#
#   scripts/ab_variants.sh build        # here (CPU, ~5 min per variant): liblbm_b200_<name>.so next to the default
#                                       # (all three variants were built once in round 1 to check that they compile
#                                       # and export the full ABI; the libraries were parked in lettuce_b200/build_<name>/,
#                                       # which does not travel to the GPU box -- rebuilding relinks the cached objects)
#   scripts/ab_variants.sh run          # on the GPU box: parity tests + config timings with every library
#
# Variants:  packed = -DLBM_KBC_PACKED=1 (KBC on FFMA2/FADD2/FMUL2)
#            spec   = -DLBM_SPECULATIVE_MASKED_LOADS=1 (label loaded together with the populations)
#            gab    = -DLBM_GENERAL_AFTER_BULK=1 (bulk kernel ignores the labels, general nodes overwrite afterwards)
set -euo pipefail
cd "$(dirname "$0")/.."
declare -A DEFINES=([packed]="-DLBM_KBC_PACKED=1" [spec]="-DLBM_SPECULATIVE_MASKED_LOADS=1"
                    [gab]="-DLBM_GENERAL_AFTER_BULK=1")

case "${1:-}" in
build)
    for name in "${!DEFINES[@]}"; do
        LBM_B200_NVCC_DEFINES="${DEFINES[$name]}" LBM_B200_BUILD_SUFFIX="$name" python -m lettuce_b200.build
    done
    ;;
run)
    mkdir -p gpurun_out
    for name in default "${!DEFINES[@]}"; do
        lib="$PWD/lettuce_b200/liblbm_b200${name:+_$name}.so"
        [ "$name" = default ] && lib="$PWD/lettuce_b200/liblbm_b200.so"
        [ -f "$lib" ] || { echo "missing $lib (run: scripts/ab_variants.sh build)"; continue; }
        echo "== $name"
        LBM_B200_LIB="$lib" python -m pytest tests -m gpu -q -x 2>&1 | tail -2
        LBM_B200_LIB="$lib" python scripts/bench_configs.py c3 c4 c5 --small | tee "gpurun_out/ab_${name}.jsonl" \
            | python -c 'import sys, json
for line in sys.stdin:
    r = json.loads(line)
    print("  %-48s %-15s %9.0f MLUPS  %.3f of peak" % (r["case"], r["streaming"], r["mlups"], r["frac_of_measured_peak"]))'
    done
    ;;
*)
    echo "usage: $0 build|run"; exit 2
    ;;
esac
