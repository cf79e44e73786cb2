/* oracle/ref_einspline.c -- TEST INFRASTRUCTURE.  Calls the reference's own einspline coefficient solve
 * (ref: src/einspline/bspline_create.c:1311-1390 create_UBspline_3d_d), compiled from /root/reference by
 * oracle/Makefile `ref`; used to pin the restated periodic solver in qmc_oracle.hpp. */
#include "einspline/bspline_base.h"
#include "einspline/bspline_structs.h"
#include "einspline/bspline_create.h"
#include <string.h>
#include <stdlib.h>
int orc_ref_einspline_create_3d_d(const int* M, const double* data, double* coefs)
{
  Ugrid g[3];
  BCtype_d bc[3];
  for (int d = 0; d < 3; ++d)
  {
    g[d].start = 0.0;
    g[d].end   = 1.0;
    g[d].num   = M[d];
    bc[d].lCode = PERIODIC;
    bc[d].rCode = PERIODIC;
    bc[d].lVal = bc[d].rVal = 0.0;
  }
  UBspline_3d_d* s = create_UBspline_3d_d(g[0], g[1], g[2], bc[0], bc[1], bc[2], (double*)data);
  size_t n = (size_t)(M[0] + 3) * (M[1] + 3) * (M[2] + 3);
  memcpy(coefs, s->coefs, n * sizeof(double));
  free(s->coefs);
  free(s);
  return 0;
}
