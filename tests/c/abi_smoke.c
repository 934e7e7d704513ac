/* abi_smoke.c — the C ABI used from plain C99 (what a cgo / JNI / ctypes binding does): one costmap, a few
 * poses, verdicts and costs checked against hand-computable answers.  Exit code 0 = ok. */
#include "cover_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                       \
    }                                                                 \
  } while (0)

int main(void) {
  cover_ctx* ctx = NULL;
  if (cover_ctx_create(0, &ctx) != COVER_OK) {
    fprintf(stderr, "no CUDA device: the library has no CPU fallback\n");
    return 2;
  }
  CHECK(cover_abi_version() == COVER_B200_ABI_VERSION);

  /* 100 x 80 map at 0.1 m, every cell costs 7, one lethal cell at (50, 40) */
  enum { SX = 100, SY = 80 };
  static uint8_t bytes[SX * SY];
  memset(bytes, 7, sizeof bytes);
  bytes[40 * SX + 50] = 254;
  cover_map* map = NULL;
  CHECK(cover_map_create(ctx, SX, SY, 0.0, 0, 0, &map) == COVER_E_INVALID_ARG); /* resolution must be positive */
  CHECK(strlen(cover_last_error(ctx)) > 0);
  CHECK(cover_map_create(ctx, SX, SY, 0.1, 0.0, 0.0, &map) == COVER_OK);
  CHECK(cover_map_upload(map, bytes, COVER_MEM_HOST) == COVER_OK);

  /* 1.0 x 0.6 m box, column-major x0,y0,x1,y1,... ; poses = {cos, sin, tx, ty} */
  const double box[8] = {0.55, 0.35, 0.55, -0.25, -0.45, -0.25, -0.45, 0.35};
  const double poses[3 * 4] = {1, 0, 2.0, 2.0, /* free */ 1, 0, 5.0, 4.0, /* covers the lethal cell */ 1, 0, 9.8, 4.0 /* leaves the map */};
  uint8_t is_free[3], status[3];
  double cost[3];
  cover_poses p;
  memset(&p, 0, sizeof p);
  p.format = COVER_POSE_CS, p.mem = COVER_MEM_HOST, p.count = 3, p.data = poses;
  cover_results out;
  memset(&out, 0, sizeof out);
  out.mem = COVER_MEM_HOST, out.is_free = is_free, out.cost = cost, out.status = status;
  CHECK(cover_area_accumulate_cost(ctx, map, box, 4, &p, &out) == COVER_OK);
  /* cells x: trunc(15.5)..trunc(25.5) = 15..25 (11), y: trunc(17.5)..trunc(23.5) = 17..23 (7) -> 77 cells x 7 */
  CHECK(status[0] == COVER_STATUS_OK && is_free[0] == 1 && cost[0] == 77 * 7);
  CHECK(is_free[1] == 0 && cost[1] == -1.0);
  CHECK(is_free[2] == 0 && cost[2] == -3.0);
  uint32_t bits[1] = {0xFFFFFFFFu};
  int64_t not_ok = -1;
  memset(&out, 0, sizeof out);
  out.mem = COVER_MEM_HOST, out.free_bits = bits, out.not_ok = &not_ok;
  CHECK(cover_area_is_free(ctx, map, box, 4, &p, &out) == COVER_OK);
  CHECK(bits[0] == 1u && not_ok == 0);
  CHECK(cover_ctx_launch_count(ctx) > 0);
  cover_map_destroy(map);
  cover_ctx_destroy(ctx);
  puts("abi_smoke: ok");
  return 0;
}
