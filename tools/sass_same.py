"""Developer check: are two objects' kernels instruction-identical?  python tools/sass_same.py a.o b.o
Used to show that a build with only developer macros / dead branches changed ships the same SASS as a tested build."""
import collections, re, subprocess, sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    ks, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = re.sub(r"_(GLOBAL__N_|INTERNAL)_[0-9a-f]+_", r"_\1_X_", m.group(1))
            cur = ks.setdefault(name, [])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append(m.group(1).strip())
    return ks


a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
diff = [k for k in sorted(set(a) | set(b)) if a.get(k) != b.get(k)]
print(f"{len(a)} / {len(b)} kernels, {len(diff)} differ")
for k in diff:
    print("  ", k[-90:], len(a.get(k, [])), len(b.get(k, [])))
sys.exit(1 if diff else 0)
