#!/bin/bash
# gpurun_retry.sh <log> <timeout> <command...>: retries while the pod answers "transient" / busy (exit code 3), at most 12 times
LOG=$1; shift; TMO=$1; shift
# the snapshot ships the in-tree .so: make sure it is built from the current sources
python -m tetwild_b200.build > /dev/null || exit 9
python -c "import oracle; oracle.build()" || exit 9
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if ! grep -q "status=transient" $LOG && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
