#!/usr/bin/env python3
"""Writes a C translation unit to stdout that holds the text of <dir>/urdf_filter.vert and .frag as two byte arrays (empty
when the files are absent): `python3 embed_shaders.py <dir> | cc -x c -c - -o ref_shaders.o`.  The reference's sources are
never written into the repository; they only travel inside the built binary under oracle/_ref/ (git-ignored)."""
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else ""
for name, sym in (("urdf_filter.vert", "ruf_ref_vert_src"), ("urdf_filter.frag", "ruf_ref_frag_src")):
    path = os.path.join(d, name)
    data = open(path, "rb").read() if os.path.exists(path) else b""
    print("const char %s[] = {%s0};" % (sym, "".join("%d," % b for b in data)))
