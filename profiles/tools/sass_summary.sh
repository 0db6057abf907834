#!/bin/bash
# cuobjdump -sass evidence for the bulk-copy flush (UBLKCP.G.S + FENCE.VIEW.ASYNC.S) and the static instruction mix
#   profiles/tools/sass_summary.sh > profiles/r2/sass_summary.txt
cd "$(dirname "$0")/../.."
LIB=ciclope_b200/libciclope_b200.so
for f in _Z16k_emit_elements2ILb0ELb0EEv14ElemEmitParams _Z16k_emit_elements2ILb0ELb1EEv14ElemEmitParams _Z13k_emit_nodes414NodeEmitParams; do
  echo "== $(echo $f | c++filt) (cuobjdump -sass -fun $f $LIB, sm_100a) =="
  cuobjdump -sass -fun $f $LIB 2>/dev/null > /tmp/_sass.txt
  grep -E 'UBLKCP|FENCE.VIEW.ASYNC|REDUX|ATOMG|VOTE' /tmp/_sass.txt | sed 's/  */ /g' | sort | uniq -c | sort -rn | head -12
  python3 - <<'P'
import re, collections
c = collections.Counter()
for line in open('/tmp/_sass.txt'):
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m: c[m.group(1)] += 1
print("static instruction mix (%d instructions): %s" % (sum(c.values()), ", ".join("%s %d" % kv for kv in c.most_common(24))))
P
  echo
done
