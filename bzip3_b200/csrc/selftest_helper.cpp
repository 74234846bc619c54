// selftest_helper.cpp -> bzip3_b200/bz3_selftest.  Spawned by libbzip3_b200.so at the first bz3_new() of a process
// (kernel_autoselect in bz3_api.cu): loads the library, lets it test the newer kernels against the proven ones on the
// given device IN THIS PROCESS and prints the result.  Whatever goes wrong here -- a hung kernel, a device fault --
// stays here; the caller just keeps the proven kernels.
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    if (argc < 3) return 64;
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) return 65;
    typedef int (*fn_t)(int, int*, int*, int*);
    fn_t fn = reinterpret_cast<fn_t>(dlsym(h, "bz3_b200_selftest"));
    if (!fn) return 66;
    int e = 0, d = 0, l = 3;
    const int rc = fn(atoi(argv[2]), &e, &d, &l);
    if (rc != 0) return rc;
    printf("BZ3SELFTEST %d %d %d\n", e, d, l);
    fflush(stdout);
    return 0;
}
