#!/usr/bin/env python3
"""Build the reference host with the libgpusim binding applied.

The reference's C host and CLI stay exactly as they are; only the sample loop +
output formatting + fwrite (gpssim.c:2190-2288) is replaced by GPUSIM_HOOK_EPOCH()
(see gpusim_hook.h and INTEGRATION.md).  This script applies that edit to a
scratch copy of /root/reference/gpssim.c at build time - reference sources are
never copied into the tracked tree - and links the result against libgpusim.so:

    integration/_build/gps-sdr-sim-gpu-int    host built with gpssim.h:4 disabled
    integration/_build/gps-sdr-sim-gpu-float  host as shipped (FLOAT_CARR_PHASE)

Both take the reference's argv unchanged.  `_build/` is git-ignored but travels to
the GPU box with gpurun.  If /root/reference is absent (GPU box) the script is a
no-op and the prebuilt binaries are used.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GPSSIM_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_build")
LIBDIR = os.path.join(ROOT, "gps_sdr_sim_b200")
# the reference Makefile's flags (Makefile:8,12); no -march, no -ffast-math.  -ffp-contract=off states what
# those flags mean on x86-64 (no FMA without -march): libgpusim reproduces "code_phase += f_code*delt" and
# "carr_phase += f_carr*delt" as a rounded product followed by a rounded sum (gpssim.c:2212, :2245).  On a
# target whose baseline has FMA (aarch64) GCC's default -ffp-contract=fast would fuse them and change
# the bytes of the reference itself; the binding is defined against the unfused arithmetic.
CFLAGS = ["-O3", "-Wall", "-D_FILE_OFFSET_BITS=64", "-ffp-contract=off"]


def apply_binding(src: str) -> str:
    """The whole edit to gpssim.c, located by the reference's own statements."""
    # (1) declare the binding after the reference's own header
    inc = '#include "gpssim.h"'
    assert src.count(inc) == 1
    src = src.replace(inc, inc + '\n#include "gpusim_hook.h"')

    # (2) replace the sample loop, the SC01/SC08/SC16 conversion and the fwrite
    #     (gpssim.c:2190-2288) by one call
    start = src.index("for (isamp=0; isamp<iq_buff_size; isamp++)")
    start = src.rindex("\n", 0, start) + 1
    last_fwrite = "fwrite(iq_buff, 2, 2*iq_buff_size, fp);"
    assert src.count(last_fwrite) == 1
    end = src.index(last_fwrite, start)
    end = src.index("}", end) + 1          # closes the "else // data_format==SC16" block
    body = src[start:end]
    # sanity: the region is the hot path and nothing else
    assert "carr_phase" in body and "iq8_buff" in body and "generateNavMsg" not in body
    src = src[:start] + "\t\tGPUSIM_HOOK_EPOCH();\n" + src[end:]

    # (2b) optional: the computeRange() pair of the epoch loop (gpssim.c:2165-2168) becomes a look-up
    #      into ranges the hook computed ahead of time with the same function
    loop = src.index("for (iumd=1; iumd<numd; iumd++)")
    m = re.compile(r"if \(!staticLocationMode\)\s*\n\s*computeRange\(&rho, eph\[ieph\]\[sv\], &ionoutc, grx, xyz\[iumd\]\);\s*\n"
                   r"\s*else\s*\n\s*computeRange\(&rho, eph\[ieph\]\[sv\], &ionoutc, grx, xyz\[0\]\);").search(src, loop)
    assert m and m.start() - loop < 1000, "computeRange pair of the epoch loop not found"
    src = src[:m.start()] + "GPUSIM_HOOK_RANGE();" + src[m.end():]

    # (3) open before the loop's clock starts, close before it stops
    m = re.search(r"^\s*tstart = clock\(\);", src, re.M)
    assert m
    src = src[:m.start()] + "\n\tGPUSIM_HOOK_OPEN();\n" + src[m.start():]
    m = re.search(r"^\s*tend = clock\(\);", src, re.M)
    assert m
    src = src[:m.start()] + "\n\tGPUSIM_HOOK_CLOSE();\n" + src[m.start():]
    return src


def build(variant: str, tmp: str) -> str:
    work = os.path.join(tmp, variant)
    os.makedirs(work)
    with open(os.path.join(REF, "gpssim.h")) as f:
        hdr = f.read()
    if variant == "int":
        hdr, n = re.subn(r"^#define FLOAT_CARR_PHASE", "// #define FLOAT_CARR_PHASE", hdr, flags=re.M)
        assert n == 1
    with open(os.path.join(work, "gpssim.h"), "w") as f:
        f.write(hdr)
    with open(os.path.join(REF, "gpssim.c")) as f:
        src = apply_binding(f.read())
    with open(os.path.join(work, "gpssim_gpu.c"), "w") as f:
        f.write(src)
    exe = os.path.join(OUT, f"gps-sdr-sim-gpu-{variant}")
    cmd = ["gcc", *CFLAGS, "-I", work, "-I", HERE, "-I", os.path.join(ROOT, "include"),
           os.path.join(work, "gpssim_gpu.c"), os.path.join(HERE, "gpusim_hook.c"),
           "-L", LIBDIR, "-lgpusim", "-Wl,-rpath,$ORIGIN/../../gps_sdr_sim_b200",
           "-lm", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit(f"build_host.py: gcc failed for variant {variant}")
    return exe


def main() -> int:
    if not os.path.isfile(os.path.join(REF, "gpssim.c")):
        print(f"build_host.py: {REF}/gpssim.c not found - keeping prebuilt integration/_build")
        return 0
    if not os.path.isfile(os.path.join(LIBDIR, "libgpusim.so")):
        raise SystemExit("build_host.py: build gps_sdr_sim_b200/libgpusim.so first")
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="gpusim_host_")
    try:
        for v in ("int", "float"):
            print("build_host.py: built", build(v, tmp))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
