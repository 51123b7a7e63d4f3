#!/bin/bash
# usage: tools/build_variant.sh NAME [-DMACRO ...]   -> build/lib_NAME.so
# Experimental builds of the whole library with extra macros; POYB200_SO=build/lib_NAME.so makes the Python binding load one.
name=$1; shift
mkdir -p build
python - "$name" "$@" <<'PY'
import sys
sys.path.insert(0, ".")
from poyd_b200 import build
print(build.build(defines=sys.argv[2:], out=f"build/lib_{sys.argv[1]}.so"))
PY
