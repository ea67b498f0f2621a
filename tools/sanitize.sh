#!/bin/bash
# compute-sanitizer targets for the wavefront kernels (the queue counters are atomics, the stack tops and ray-space
# rows live in shared memory).  Usage, on a box whose driver allows the sanitizer (this pool's boxes refuse it, see
# gpurun_out/memcheck.log of round 1):   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
# Runs the smoke frame (Cornell box, 480x270, three foveation passes, depth 4) and one subframe batch under the tool.
TOOL=${1:-memcheck}
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/compute-sanitizer --tool "$TOOL" --error-exitcode 1 --kernel-regex kns=fovpt \
  python -c "import __graft_entry__ as g; g.smoke()"
