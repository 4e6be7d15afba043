"""Build glue/duckdb_gpu_glue.c against the mock DuckDB C API (glue/mock/) into glue/_build/: what the tests load.
With a real DuckDB installation: replace -Iglue/mock by the directory of the real duckdb.h and -lduckdb_mock by -lduckdb,
and drop -DDMB_GLUE_STANDALONE when linking next to duckdb_native.c."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
MOCK = os.path.join(OUT, "libduckdb_mock.so")
GLUE = os.path.join(OUT, "libduckdb_gpu_glue.so")
LIBDIR = os.path.join(ROOT, "duckdb.mbt_b200", "csrc")


def _newer(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def build(force: bool = False):
    os.makedirs(OUT, exist_ok=True)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    mock_src = os.path.join(HERE, "mock", "libduckdb_mock.c")
    hdrs = [os.path.join(HERE, "mock", "duckdb.h"), os.path.join(ROOT, "include", "duckdb_mb_gpu.h"), os.path.join(ROOT, "include", "moonbit_standin.h")]
    if force or _newer(MOCK, [mock_src] + hdrs):
        subprocess.check_call([cc, "-O1", "-g", "-Wall", "-Wextra", "-fPIC", "-shared", "-I" + os.path.join(HERE, "mock"), "-o", MOCK, mock_src])
    glue_src = os.path.join(HERE, "duckdb_gpu_glue.c")
    gpu_lib = os.path.join(LIBDIR, "libduckdb_mb_gpu.so")
    if force or _newer(GLUE, [glue_src, gpu_lib, MOCK] + hdrs):
        subprocess.check_call([cc, "-O1", "-g", "-Wall", "-Wextra", "-std=gnu11", "-fPIC", "-shared", "-DDMB_GLUE_STANDALONE",
                               "-I" + os.path.join(HERE, "mock"), "-I" + os.path.join(ROOT, "include"), "-o", GLUE, glue_src,
                               "-L" + LIBDIR, "-L" + OUT, "-lduckdb_mb_gpu", "-lduckdb_mock",
                               "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + OUT])
    return MOCK, GLUE


if __name__ == "__main__":
    print(build(force=True))
