#!/bin/bash
# Local helper (runs in the build container, not on the GPU box): retries `gpurun` while the pod answers
# "busy / draining" (status "transient", nothing charged).  Usage: tools/gpu/gr.sh [gpurun options] -- '<command>'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
