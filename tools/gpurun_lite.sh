#!/bin/bash
# Dev helper: run a gpurun command WITHOUT shipping the ~170 MB of tensor-product spec libraries
# (GEMM / runtime-only experiments).  Restores .gpurunignore afterwards.
cd "$(dirname "$0")/.."
cp .gpurunignore /tmp/.gpurunignore.bak 2>/dev/null || : > /tmp/.gpurunignore.bak
trap 'if [ -s /tmp/.gpurunignore.bak ]; then cp /tmp/.gpurunignore.bak .gpurunignore; else rm -f .gpurunignore; fi' EXIT
echo "nequip_b200/lib/nqbspec_*" >> .gpurunignore
/usr/local/graft/bin/gpurun "$@"
