"""profiles/*_sass_excerpt.txt: SASS mnemonic counts per kernel of the built library.

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt

cuobjdump -sass of picasso_b200/libpicasso_b200.so (sm_100a cubins), per kernel the number of
instructions of the mnemonics that show what the kernel is built from: UBLKCP (1-D bulk async copy, TMA
engine), SYNCS (mbarrier), LDGSTS (cp.async), VIMNMX*.U16x2 (packed uint16 min/max), ATOMS.CAST.SPIN
(shared-memory float atomic), DFMA / FFMA (which pipe the arithmetic runs on), MUFU.* -- and, as the
negative evidence the spec asks about, UTMALDG / UTCMMA / LDTM (tensor-map TMA / tcgen05), which do not
occur: no kernel here is a dense contraction."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "picasso_b200", "libpicasso_b200.so")
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "VIMNMX3.U16x2", "VIMNMX.U16x2", "ATOMS.CAST.SPIN", "ATOMS.ADD", "REDG",
         "DFMA", "FFMA", "MUFU.RCP64H", "MUFU.RCP", "MUFU.EX2", "F2F", "UTMALDG", "UTCMMA", "UTCHMMA", "LDTM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                           capture_output=True, text=True).stdout.splitlines()
    blocks = re.split(r"\s+Function : \S+\n", sass)[1:]
    print("# SASS mnemonic counts per kernel of picasso_b200/libpicasso_b200.so (cuobjdump -sass, sm_100a;")
    print("# tools/sass_excerpt.py).  Total instructions first; mnemonics that do not occur are omitted.")
    tot = collections.Counter()
    for name, body in zip(names, blocks):
        ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.x]*)", body)
        c = collections.Counter()
        for op in ops:
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    c[w] += 1
                    break
        tot.update(c)
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        print(f"{short}\n    {len(ops)} instructions: " + ", ".join(f"{w} x{c[w]}" for w in WATCH if c[w]))
    print("# library total: " + ", ".join(f"{w} x{tot[w]}" for w in WATCH))


if __name__ == "__main__":
    main()
