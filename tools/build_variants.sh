#!/bin/bash
# development: build A/B variants of the library into build/rt_<name>.so.   usage: tools/build_variants.sh "name:-DX=1 -DY=2" ...
cd "$(dirname "$0")/../cs397raytracingsp22_b200" && mkdir -p ../build
for v in "$@"; do
  n=${v%%:*}; d=${v#*:}
  python build.py $d --out=../build/rt_$n.so > /dev/null 2>&1 && echo "built $n ($d)" || echo "FAILED $n"
done
