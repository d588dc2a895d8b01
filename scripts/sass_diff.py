"""Which kernels changed? Per-kernel SASS digest of two object files / libraries (cuobjdump, no GPU needed):

    python scripts/sass_diff.py old.o new.o

Prints SAME / DIFF / NEW / GONE per demangled kernel name (anonymous-namespace hashes normalised). Used before committing a
change made without GPU access: kernels of the validated default path must come out SAME."""
import subprocess, re, sys, hashlib
def funcs(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    res, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name: res[name] = body
            name, body = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            body.append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    if name: res[name] = body
    dem = {}
    for k, v in res.items():
        d = subprocess.run(["cu++filt", k], stdout=subprocess.PIPE, text=True).stdout.strip()
        d = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+_cu_[0-9a-f]+", "ANON", d)
        dem[d] = hashlib.sha1("\n".join(v).encode()).hexdigest()[:12] + f" ({len(v)} instr)"
    return dem
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
for k in sorted(set(a) | set(b)):
    st = "SAME" if a.get(k) == b.get(k) else ("NEW" if k not in a else "GONE" if k not in b else "DIFF")
    print(st, k[:110], a.get(k, ""), b.get(k, "") if st != "SAME" else "")
