// placeholder until the bulk-async kernel lands
#include "cdm_internal.hpp"
int cdm_k_apply_p3(cdm_op *op, const int32_t *, const double *, double *)
{ return cdm_fail(op->sp->ctx, CDM_EUNSUP, "p3 kernel not built"); }
