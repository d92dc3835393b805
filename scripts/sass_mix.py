"""Static SASS opcode mix per kernel of libqaoa_b200.so (straight-line pass programs: static ~ per-tile dynamic).
usage: cuobjdump -sass qaoa_b200/libqaoa_b200.so > /tmp/sass.txt; python scripts/sass_mix.py /tmp/sass.txt [filter]"""
import collections, re, sys
cur = None
mix = collections.OrderedDict()
for line in open(sys.argv[1]):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); mix[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        mix[cur][m.group(2)] += 1
flt = sys.argv[2] if len(sys.argv) > 2 else "tile_prog_kernelI6float2"
for k, c in mix.items():
    if flt not in k: continue
    tot = sum(c.values())
    print("%s\n   total %d | %s" % (k[:70], tot, " ".join("%s=%d" % kv for kv in c.most_common(24))))
