#!/bin/bash
for v in "" _vsl3 _vsl4; do
  EVPLP_LIB=$PWD/evplp_b200/lib/libevplp_b200$v.so python scripts/config_runs.py C3s 2>&1 | tail -1 | cut -c80-330
done
