/* The block-level C-ABI from plain C: Encoder_Block.forward (layers.py:174-193) as ONE call.
 *
 * Runs WITHOUT a GPU: the library's dry-run trace (dg_debug_trace) makes every kernel entry point record its name and
 * arguments instead of launching, so the program below prints the ten launches dg_block_fwd would issue for the (fake) buffer
 * addresses it is given.  On a B200, drop the two dg_debug_trace* calls, pass device pointers and a stream.
 *
 *   gcc -std=c99 -Iinclude examples/block_fwd_trace.c -Ldruggen_b200 -ldruggen_b200 -Wl,-rpath,$PWD/druggen_b200 -o /tmp/block_fwd_trace
 */
#include <stdio.h>
#include <stdint.h>
#include "druggen_b200.h"

int main(void) {
  enum { B = 4, N = 45, D = 128, H = 384, HEADS = 8 };
  void* io[DG_BLK_COUNT] = {0};
  const float* params[DG_BLOCK_PARAMS];
  static char text[1 << 16];
  int i;
  /* caller-owned buffers: here just distinct addresses, 256 MB apart */
  const int used[] = {DG_BLK_X, DG_BLK_Y, DG_BLK_X_OUT, DG_BLK_Y_OUT, DG_BLK_X1, DG_BLK_Q, DG_BLK_K, DG_BLK_V, DG_BLK_G, DG_BLK_ON,
                      DG_BLK_X3, DG_BLK_Y3, DG_BLK_A16};
  for (i = 0; i < (int)(sizeof used / sizeof used[0]); ++i) io[used[i]] = (void*)(uintptr_t)(0x1000000000ull + 0x10000000ull * used[i]);
  for (i = 0; i < DG_BLOCK_PARAMS; ++i) params[i] = (const float*)(uintptr_t)(0x2000000000ull + 0x10000000ull * i);
  dg_debug_trace(1);
  if (dg_block_fwd(io, params, B, N, D, H, HEADS, DG_BLKF_EDGE_OUT, 1e-5f, (void*)(uintptr_t)0x4000000000ull, 2 * (H / 128) * 32768, NULL)) {
    fprintf(stderr, "dg_block_fwd rejected: %s\n", dg_last_error());
    return 1;
  }
  dg_debug_trace_read(text, sizeof text);
  dg_debug_trace(0);
  fputs(text, stdout);
  printf("abi %d, %lld launches issued by the block-level entry points\n", dg_abi_version(), dg_native_launches());
  return 0;
}
