#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: INTEGRATION.md section B made executable.

Takes the UNMODIFIED reference sources where they lie (/root/reference/src), applies -- in a temporary
directory that is deleted afterwards -- exactly the binding INTEGRATION.md shows a maintainer of the
reference (the reference's own hook.c / client.c keep everything else: interposition, the gate, the
protocol, the threads), compiles it with the reference's flags and links it against our C-ABI library:

    oracle/_ref/libnvshare_bound.so  =  reference hook + reference client  +  libnvs_engine.so

Nothing of the reference is copied into the repository: the edits below name short anchor strings of the
reference's code (file:line in the comments) and the only output is the shared object in oracle/_ref/.
tests/test_bound_reference.py runs two oversubscribed clients under the REFERENCE daemon with it.

    python oracle/bind_reference.py [reference root] [output .so]
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("/root/reference")
OUT = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "oracle" / "_ref" / "libnvshare_bound.so"
BUILD = ROOT / "nvshare_b200" / "_build"

HOOK_STATE = r'''
/* ---- binding to libnvs_engine.so (INTEGRATION.md B) ---- */
#include "nvshare_engine.h"
nvs_engine *nvshare_engine;                       /* created lazily: needs the app's context; client.c uses it too */
static void *nvs_cuda_handle;
static void *nvs_resolve(const char *s) { return real_dlsym_225(nvs_cuda_handle, s); }
static nvs_engine *engine_get(void)
{
	if (!nvshare_engine) {
		nvs_engine_config cfg;
		nvs_engine_default_config(&cfg);
		cfg.resolve = nvs_resolve;                /* the REAL driver symbols, not the hooks */
		if (nvs_engine_create(&cfg, &nvshare_engine) != 0)
			log_fatal("swap engine");
	}
	return nvshare_engine;
}
'''


def edit(text, anchor, replacement, what, count=1, after=None):
    """Replace `anchor` (the first occurrence after `after`, when given) and insist that it was there."""
    start = 0
    if after is not None:
        start = text.find(after)
        assert start >= 0, f"{what}: context {after!r} not found in the reference source"
    at = text.find(anchor, start)
    assert at >= 0, f"{what}: anchor {anchor!r} not found in the reference source"
    return text[:at] + replacement + text[at + len(anchor):]


def bind_hook(src):
    # src/hook.c:50 -- file-scope state, right after the forward declaration of the real dlsym
    a = "static void *real_dlsym_225(void *handle, const char *symbol);"
    src = edit(src, a, a + HOOK_STATE, "engine state")
    # src/hook.c:145 -- remember the driver's handle for the resolver
    a = 'cuda_handle = dlopen("libcuda.so", RTLD_LAZY);'
    src = edit(src, a, a + "\n\tnvs_cuda_handle = cuda_handle;", "driver handle")
    # src/hook.c:673 -- cuMemAlloc: the allocation comes from the engine
    a = "result = real_cuMemAllocManaged(dptr, bytesize, CU_MEM_ATTACH_GLOBAL);"
    src = edit(src, a, "{ uint64_t p_ = 0; result = (CUresult)nvs_alloc(engine_get(), &p_, bytesize); "
                       "if (result == CUDA_SUCCESS) *dptr = (CUdeviceptr)p_; }", "cuMemAlloc")
    # src/hook.c:691 -- cuMemFree: ours first, the driver's for everything else
    a = "result = real_cuMemFree(dptr);"
    src = edit(src, a, "result = nvshare_engine ? (CUresult)nvs_free(nvshare_engine, dptr) : (CUresult)NVS_E_NOT_OURS;\n"
                       "\tif ((int)result == NVS_E_NOT_OURS) result = real_cuMemFree(dptr);", "cuMemFree")
    return src


def bind_client(src):
    # file scope (src/client.c:40)
    a = "pthread_t client_tid;"
    src = edit(src, a, '#include "nvshare_engine.h"\nextern nvs_engine *nvshare_engine;\n' + a, "client state")
    # src/client.c:298-304 -- LOCK_OK: fetch BEFORE the gate opens (VMM memory cannot fault)
    a = "need_lock = 0;"
    src = edit(src, a, "if (nvshare_engine) {\n\t\t\t\tif (nvs_fetch_all(nvshare_engine, NULL) != 0) log_fatal(\"fetch\");\n"
                       "\t\t\t\tnvs_set_resident_mode(nvshare_engine, 1);\n\t\t\t}\n\t\t\t" + a, "LOCK_OK", after="case LOCK_OK:")
    # src/client.c:308-317 -- DROP_LOCK: evict after the context is drained and LOCK_RELEASED is out
    a = 'log_debug("Sent %s", message_type_string[out_msg.type]);'
    src = edit(src, a, a + "\n\t\t\t\tif (nvshare_engine) {\n\t\t\t\t\tnvs_set_resident_mode(nvshare_engine, 0);\n"
                           "\t\t\t\t\tif (nvs_evict(nvshare_engine, 0, NULL) != 0) log_fatal(\"evict\");\n\t\t\t\t}",
               "DROP_LOCK", after="case DROP_LOCK:")
    # src/client.c:472-476 -- early release
    a = 'log_debug("Sent %s", message_type_string[release_msg.type]);'
    src = edit(src, a, a + "\n\t\t\tif (nvshare_engine) {\n\t\t\t\tcuda_sync_context();\n\t\t\t\tnvs_set_resident_mode(nvshare_engine, 0);\n"
                           "\t\t\t\tif (nvs_evict(nvshare_engine, 0, NULL) != 0) log_fatal(\"evict\");\n\t\t\t}", "early release")
    return src


def main():
    src = REF / "src"
    if not (src / "hook.c").exists():
        print(f"reference sources not present ({src}): nothing to bind")
        return 0
    if not (BUILD / "libnvs_engine.so").exists():
        print(f"{BUILD}/libnvs_engine.so is missing: build the product first", file=sys.stderr)
        return 1
    tmp = Path(tempfile.mkdtemp(prefix="nvs_bind_"))
    try:
        (tmp / "hook.c").write_text(bind_hook((src / "hook.c").read_text()))
        (tmp / "client.c").write_text(bind_client((src / "client.c").read_text()))
        OUT.parent.mkdir(parents=True, exist_ok=True)
        cmd = ["gcc", "-O2", "-g", "-fPIC", "-w", "-D_GNU_SOURCE", f"-I{src}", f"-I{ROOT / 'include'}",
               str(tmp / "hook.c"), str(tmp / "client.c"), str(src / "common.c"), str(src / "comm.c"),
               "-shared", "-Wl,-soname=libnvshare.so", f"-Wl,--version-script={src / 'libnvshare-symbols.ld'}",
               "-Wl,--exclude-libs,ALL", f"-L{BUILD}", "-lnvs_engine", "-Wl,-rpath,$ORIGIN/../../nvshare_b200/_build",
               "-o", str(OUT), "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr, file=sys.stderr)
            return r.returncode
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(f"bound reference library: {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
