#!/bin/bash
# A/B of an environment switch: tools/ab_env.sh VAR val1 val2 ...   (quick decoder timing + phases per value)
var=$1; shift
for v in "$@"; do export $var=$v; echo "=== $var=$v"; bash tools/quick_dec.sh ab_$v | grep -v "^cluster B="; done
