"""Embed kernels/*.cuh as C++ raw string literals -> kernels_embed.inc (build step)."""
import glob, os, sys
here = os.path.dirname(os.path.abspath(__file__))
out = []
for path in sorted(glob.glob(os.path.join(here, "kernels", "*.cuh"))):
    name = os.path.basename(path)
    txt = open(path).read()
    assert ')B2EMBED"' not in txt
    # MSVC-free toolchain: a single raw literal per header is fine for gcc
    out.append('{"%s", R"B2EMBED(%s)B2EMBED"},' % (name, txt))
open(os.path.join(here, "kernels_embed.inc"), "w").write("\n".join(out) + "\n")
