// TEST INFRASTRUCTURE ONLY: sequential host build of the cleaning kernel logic (clean_core.cuh with
// -DMTB_HOST_EMUL) so tests can compare it with cv2 / the reference in a container that has no GPU.
// Never linked into libmtb200.so; the product has no CPU path.
#define MTB_HOST_EMUL 1
#include "../../mangatranslator_b200/csrc/clean_core.cuh"

extern "C" {
unsigned long long emul_workspace_words(int cw, int ch, int max_runs) {
  return mtbclean::workspace_words(cw, ch, max_runs);
}
int emul_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(mtbclean::Params);
    case 1: return sizeof(mtbclean::Job);
    case 2: return sizeof(mtbclean::Result);
  }
  return -1;
}
void emul_clean_job(const mtbclean::Params* P, const mtbclean::Job* J, mtbclean::Result* R) {
  mtbclean::Shared sh;
  memset(&sh, 0, sizeof(sh));
  mtbclean::clean_job(*P, *J, *R, &sh);
}
int emul_otsu(const unsigned int* hist, unsigned int total) { return mtbclean::otsu_threshold(hist, total); }
int emul_hsv_sat(int b, int g, int r) { return mtbclean::hsv_saturation(b, g, r); }
}
