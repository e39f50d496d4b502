/* Plain C against the C ABI (include/ifx_b200.h): step a handful of positions through the rules kernel with the
 * host-buffer entry point, the way a non-Python host program would.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -o c_abi_demo -ldl && ./c_abi_demo alpha-zero-general-inflexion_b200/libifx_b200.so
 *
 * Needs a CUDA device at run time (the library has no CPU path); it only has to COMPILE in the build container,
 * which tests/test_cabi.py checks. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "ifx_b200.h"

typedef int (*mask_step_host_fn)(const ifx_state*, const int32_t*, int64_t, uint64_t*, ifx_state*, int32_t*);

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "alpha-zero-general-inflexion_b200/libifx_b200.so";
  void* lib = dlopen(path, RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  mask_step_host_fn step;
  int (*version)(void);
  *(void**)(&step) = dlsym(lib, "ifx_rules_mask_step_host");      /* the POSIX idiom for object -> function pointer */
  *(void**)(&version) = dlsym(lib, "ifx_version");
  if (!step || !version) { fprintf(stderr, "missing symbols\n"); return 2; }

  /* the start position (empty 7x7 board, RED to move, max_turns 343) and a spawn on cell (3,3): action 6*49 + 3*7 + 3 */
  ifx_state in[2], out[2];
  int32_t actions[2] = {6 * 49 + 3 * 7 + 3, 0 /* a spread from an empty cell: invalid */}, status[2];
  uint64_t masks[2][7];
  memset(in, 0, sizeof(in));
  in[0].max_turns = in[1].max_turns = 343;
  int rc = step(in, actions, 2, &masks[0][0], out, status);
  if (rc != IFX_OK) { fprintf(stderr, "ifx_rules_mask_step_host failed: %d\n", rc); return 1; }
  printf("libifx_b200 version %d\n", version());
  printf("start position: spawn mask %013llx (49 empty cells), spread masks %llx\n",
         (unsigned long long)masks[0][6], (unsigned long long)masks[0][0]);
  printf("after spawn at (3,3): red %013llx, turn %u, side to move %s, status %d\n",
         (unsigned long long)out[0].red, (unsigned)out[0].turn, out[0].player ? "BLUE" : "RED", status[0]);
  printf("invalid action: status %d (IFX_F_BADACTION = %d), state passed through unchanged: %s\n", status[1],
         IFX_F_BADACTION, memcmp(&in[1], &out[1], sizeof(ifx_state)) == 0 ? "yes" : "no");
  int ok = masks[0][6] == ((1ull << 49) - 1) && masks[0][0] == 0 && out[0].red == (1ull << 24) && out[0].turn == 1 &&
           out[0].player == 1 && status[0] == 0 && status[1] == IFX_F_BADACTION;
  printf(ok ? "ok\n" : "MISMATCH\n");
  dlclose(lib);
  return ok ? 0 : 1;
}
