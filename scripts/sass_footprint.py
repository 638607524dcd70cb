#!/usr/bin/env python
"""Code bytes of one kernel attributed to source lines / functions (dev tool, CPU only).

  python scripts/sass_footprint.py [kernel-substring] [--lines]

Extracts the cubin from lib/obj/libpp_b200_cabi.o, disassembles with line info (nvdisasm -g) and sums 16 bytes per SASS
instruction per (file, line); lines are mapped to the enclosing function by a brace-free heuristic (the last preceding
line that looks like a function header).  Used for the instruction-footprint work on pp_search_kernel.
"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_ROOT = os.environ.get("PP_SRC_ROOT", ROOT)      # tree the object was compiled from (line -> function mapping)
OBJ = os.path.join(ROOT, "path_planning_pkg_b200", "lib", "obj", "libpp_b200_cabi.o")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    kern = args[0] if args else "pp_search_kernel"
    obj = args[1] if len(args) > 1 else OBJ
    show_lines = "--lines" in sys.argv
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    per_line = collections.Counter()
    inside, cur = False, ("?", 0)
    total = 0
    for ln in txt:
        if ln.startswith("\t.section\t.text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.match(r"^(\$\S+):", ln)             # compiler-internal subroutine (IEEE division / sqrt / fmod slow paths): no line info
        if m:
            cur = ("<internal>", 0)
            internal_name = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            per_line[cur] += 16
            total += 16
    # map lines to functions
    headers = {}
    for f in set(k[0] for k in per_line):
        path = None
        for r, _, fs in os.walk(os.path.join(SRC_ROOT, "path_planning_pkg_b200", "csrc")):
            if f in fs:
                path = os.path.join(r, f)
        hs = []
        if path:
            for i, s in enumerate(open(path, errors="replace"), 1):
                if re.match(r"^(template\s*<.*>\s*)?(PP_HD|PP_HD_NOINLINE_FN|PP_HD_NOINLINE|__global__|__device__|static|inline|PP_DEV)\b.*\(", s) or \
                   re.match(r"^\s{4}(PP_HD|PP_HD_NOINLINE_FN)\b.*\(", s):
                    name = re.search(r"([A-Za-z_0-9]+)\s*\(", s)
                    hs.append((i, name.group(1) if name else s.strip()[:40]))
        headers[f] = hs
    per_fn = collections.Counter()
    for (f, l), b in per_line.items():
        fn = "?"
        for i, name in headers.get(f, []):
            if i <= l:
                fn = name
            else:
                break
        per_fn[(f, fn)] += b
    print(f"{kern}: {total} bytes of SASS ({total // 16} instructions)")
    for (f, fn), b in per_fn.most_common(60):
        print(f"  {b:8d}  {100.0 * b / total:5.1f} %  {f}:{fn}")
    if show_lines:
        print("top lines:")
        for (f, l), b in per_line.most_common(60):
            print(f"  {b:8d}  {f}:{l}")


if __name__ == "__main__":
    main()
