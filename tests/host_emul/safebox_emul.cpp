// TEST INFRASTRUCTURE ONLY: sequential host build of the safe-text-box kernel logic (safebox_core.cuh with
// -DMTB_HOST_EMUL) so tests can compare it with the reference / the oracle in a container that has no GPU.
// Never linked into libmtb200.so; the product has no CPU path.
#define MTB_HOST_EMUL 1
#include "../../mangatranslator_b200/csrc/safebox_core.cuh"

extern "C" {
int emul_safebox_sizeof(int which) { return which == 0 ? sizeof(mtbsafe::Job) : sizeof(mtbsafe::Result); }
// bounds pass of the device path (safebox.cu), restated sequentially: running maxima of {-x0, -y0, x1, y1}
void emul_safebox_bounds(const mtbsafe::Job* J, mtbsafe::Result* R) {
  memset(R, 0x80, sizeof(*R));
  for (int y = 0; y < J->H; ++y)
    for (int x = 0; x < J->W; ++x)
      if (J->mask[(long long)y * J->pitch + x]) {
        if (-x > R->mask_bbox[0]) R->mask_bbox[0] = -x;
        if (-y > R->mask_bbox[1]) R->mask_bbox[1] = -y;
        if (x > R->mask_bbox[2]) R->mask_bbox[2] = x;
        if (y > R->mask_bbox[3]) R->mask_bbox[3] = y;
      }
}
void emul_safebox_job(const mtbsafe::Job* J, mtbsafe::Result* R) {
  mtbsafe::Shared sh;
  memset(&sh, 0, sizeof(sh));
  emul_safebox_bounds(J, R);
  mtbsafe::safe_job(*J, *R, &sh);
}
}
