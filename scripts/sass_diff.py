"""Per-kernel SASS comparison of two builds of the library (refactors that must not change the generated code):
    cuobjdump -sass old.so > a.sass; cuobjdump -sass new.so > b.sass; python scripts/sass_diff.py a.sass b.sass
Compares every kernel's instruction stream and encodings (addresses stripped); prints the kernels that differ."""
import re
import sys


def funcs(path):
    d, cur, buf = {}, None, []
    for line in open(path):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                d[cur] = buf
            cur, buf = m.group(1), []
        elif cur is not None:
            text = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line).strip()
            if text:
                buf.append(text)
    if cur:
        d[cur] = buf
    return d


if __name__ == "__main__":
    a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
    only = set(a) ^ set(b)
    diff = [k for k in a if k in b and a[k] != b[k]]
    print(f"{len(a)} / {len(b)} kernels; only in one build: {len(only)}; SASS differs: {len(diff)}")
    for k in sorted(only) + diff:
        print("  ", k[:160])
    sys.exit(1 if (only or diff) else 0)
