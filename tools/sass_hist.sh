#!/bin/bash
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{on=(index($0,pat)>0)} { if (on && match($0, /^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//)) print $2 }' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn > /tmp/hist.out
head -${3:-16} /tmp/hist.out | paste - - - -
awk '{n+=$1} END{print "total", n}' /tmp/hist.out
