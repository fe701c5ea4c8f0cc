"""Compare the SASS instruction streams of kernel instantiations between two object files (or against a saved
normalised dump).  Used after adding template-gated experimental code paths to prove that the default
instantiations still compile to exactly the instructions that were validated on hardware.

  python tools/sass_identity.py dump lca_b200/ops/build/fmha_fwd_sm100.o > /tmp/base.sass      # at the validated commit
  python tools/sass_identity.py check /tmp/base.sass lca_b200/ops/build/fmha_fwd_sm100.o 'ELb0EEEv'   # later
  python tools/sass_identity.py hash lca_b200/ops/build/*.o > profiles/sass_hashes_r1.json             # compact baseline
  python tools/sass_identity.py checkhash profiles/sass_hashes_r1.json lca_b200/ops/build/*.o
"""
import hashlib
import json
import re
import subprocess
import sys


def dump(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            cur = line.split(":", 1)[1].strip()
            funcs[cur] = []
        elif cur and re.match(r"^\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(re.sub(r"/\*[0-9a-f]{4,}\*/|/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return funcs


def load(path):
    funcs, cur = {}, None
    for line in open(path):
        if line.startswith("Function : "):
            cur = line[len("Function : "):].strip()
            funcs[cur] = []
        elif cur:
            funcs[cur].append(line.rstrip("\n"))
    return funcs


def key(name):
    """Strip template arguments appended after validation (trailing `ELb0` flags) for matching, and the path-dependent
    hash nvcc puts into the mangled name of anonymous-namespace symbols (a checkout at another path must compare equal)."""
    name = re.sub(r"_GLOBAL__N__[0-9a-f]+_(\d+)_(\w+?)_cu_[0-9a-f]+", r"_GLOBAL__N__\1_\2_cu", name)
    return re.sub(r"(ELb0)+EEEv", "EEEv", name)


def hashes(objs):
    out = {}
    for o in objs:
        for n, body in dump(o).items():
            out[key(n)] = hashlib.sha1("\n".join(body).encode()).hexdigest()
    return out


if __name__ == "__main__":
    if sys.argv[1] == "hash":
        print(json.dumps(hashes(sys.argv[2:]), indent=0, sort_keys=True))
    elif sys.argv[1] == "checkhash":
        base, new = json.load(open(sys.argv[2])), hashes(sys.argv[3:])
        bad = [n for n, h in base.items() if new.get(n) != h]
        for n in bad:
            print("DIFFERS:", n)
        print(f"{len(base) - len(bad)}/{len(base)} validated instantiations identical; {len(set(new) - set(base))} new")
        sys.exit(1 if bad else 0)
    elif sys.argv[1] == "dump":
        for n, body in dump(sys.argv[2]).items():
            print("Function : " + n)
            print("\n".join(body))
    else:
        base, new = load(sys.argv[2]), dump(sys.argv[3])
        newk = {key(n): b for n, b in new.items() if not re.search(r"ELb1EEEv", n) or key(n) == n}
        bad = [n for n, b in base.items() if newk.get(key(n)) != b]
        print(f"{len(base) - len(bad)}/{len(base)} instantiations identical")
        sys.exit(1 if bad else 0)
