#!/bin/bash
# usage: tools/sass_lines.sh build/x.o <kernel-name-substring>   -> SASS instruction count per source line
f=$(realpath "$1"); d=$(mktemp -d); ( cd $d && cuobjdump -xelf all "$f" >/dev/null && nvdisasm -gi *.cubin 2>/dev/null ) | python3 -c "
import sys,re,collections
pat=sys.argv[1]
cnt=collections.Counter(); cur=None; fn=None; first=None
for line in sys.stdin:
    m=re.match(r'\s*\.text\.(\S+):',line)
    if m:
        fn=m.group(1)
        if pat in fn and first is None: first=fn
        continue
    m=re.search(r'//## File \"([^\"]+)\", line (\d+)',line)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/',line) and fn==first: cnt[cur]+=1
print(first)
for k,v in sorted(cnt.items(), key=lambda kv:-kv[1])[:int(sys.argv[2])]: print(v,k)
print('total',sum(cnt.values()))
" "$2" "${3:-20}"
rm -rf $d
