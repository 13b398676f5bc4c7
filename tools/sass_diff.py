"""Compare the SASS of every kernel in two build directories (object files of egot2_b200/build*):

    python tools/sass_diff.py <old_build_dir> <new_build_dir>

Used to show that a change which only ADDS translation units / template instantiations leaves every kernel that was
validated on hardware byte-identical (e.g. build the validated commit in a scratch `git worktree`, then compare its
egot2_b200/build with the current one).  Instruction addresses are stripped; anonymous-namespace hashes in mangled
names (they depend on the source path) are normalised."""
import os
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    d, cur, name = {}, [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                d[name] = "\n".join(cur)
            name, cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_[A-Za-z0-9_]+?_cu_[0-9a-f]{8}", "_ANON_", m.group(1)), []
        else:
            line = re.sub(r"/\*[0-9a-f]{4}\*/", "", line)
            if line.strip():
                cur.append(line)
    if name:
        d[name] = "\n".join(cur)
    return d


def main():
    old_dir, new_dir = sys.argv[1], sys.argv[2]
    same_n = bad_n = 0
    for o in sorted(os.listdir(old_dir)):
        if not o.endswith(".o") or not os.path.exists(os.path.join(new_dir, o)):
            continue
        a, b = kernels(os.path.join(old_dir, o)), kernels(os.path.join(new_dir, o))
        same = [k for k in a if k in b and a[k] == b[k]]
        diff = [k for k in a if k in b and a[k] != b[k]]
        gone = [k for k in a if k not in b]
        new = [k for k in b if k not in a]
        same_n += len(same)
        bad_n += len(diff) + len(gone)
        print(f"{o:22s} old {len(a):3d}  identical {len(same):3d}  changed {len(diff)}  gone {len(gone)}  new {len(new)}")
        for k in diff:
            print("    CHANGED:", k[:120])
        for k in gone:
            print("    GONE (renamed?):", k[:120])
    print("TOTAL identical", same_n, "changed/gone", bad_n)
    return 0 if bad_n == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
