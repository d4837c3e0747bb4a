"""Stage the reference's own modules for the CPU reference arm (`bench.py --impl reference`).

    python baseline/stage_reference.py          # needs /root/reference (build container only)

The reference (dongzhuoyao/uspace) is plain Python scripts: no setup.py / pyproject.toml, so it cannot be pip-installed
into baseline/_ref.  This recipe copies the IMPORT CLOSURE of the velocity-field modules - libs/uvit.py
(UViT.forward, libs/uvit.py:306-351) and libs/uvit_t2i.py - as they lie under /root/reference into baseline/_ref/,
unmodified, so that the GPU box (which has no /root/reference) can time the reference's own torch-eager CPU path.
baseline/_ref/ is git-ignored (never part of this repository's history) but not gpurun-ignored: it travels with the
snapshot like the built .so.  Nothing but bench.py's reference arm imports it.
"""
import ast
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
ROOTS = ["libs/uvit.py", "libs/uvit_t2i.py"]


def local_module(name, cur_pkg):
    """Path (relative to REF) of a reference-local module `name`, or None for third-party / stdlib modules."""
    for cand in (name.replace(".", "/") + ".py", name.replace(".", "/") + "/__init__.py"):
        if os.path.exists(os.path.join(REF, cand)):
            return cand
    return None


def closure():
    todo, seen = list(ROOTS), set()
    while todo:
        rel = todo.pop()
        if rel in seen:
            continue
        seen.add(rel)
        pkg = os.path.dirname(rel).replace("/", ".")
        init = os.path.join(os.path.dirname(rel), "__init__.py")
        if os.path.dirname(rel) and os.path.exists(os.path.join(REF, init)):
            todo.append(init)
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                base = node.module or ""
                if node.level:
                    base = ".".join([p for p in (pkg, base) if p])
                names = [base] + [base + "." + a.name for a in node.names]
            for n in names:
                m = local_module(n, pkg)
                if m:
                    todo.append(m)
    return sorted(seen)


def main():
    if not os.path.isdir(REF):
        print("stage_reference: /root/reference not present (GPU box): nothing to do")
        return 0
    files = closure()
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write("unmodified copies from /root/reference (dongzhuoyao/uspace), import closure of " + ", ".join(ROOTS) + "\n")
        f.write("\n".join(files) + "\n")
    print(f"stage_reference: {len(files)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
