/*
 * flatten.h -- AoS -> flat-array view of the reference engine's global objects.
 *
 * Compiled ONLY against the reference's own headers (-I$REF/src/solver), never copied from them:
 * this is the "reference-side binding" half of the seam.  It reads the exported
 * globals Node[], Link[], Conduit[], ... (globals.h:146-169) of the engine it is loaded into and
 * fills the plain-pointer descriptor of include/swmm_b200.h.
 */
#ifndef SWB_FLATTEN_H
#define SWB_FLATTEN_H
#include "swmm_b200.h"

typedef struct swb_flat {
    swb_network_desc desc;
    swb_options      opt;
    void           **owned;     /* every malloc'ed block, for swb_flat_free */
    int              n_owned, cap_owned;
} swb_flat;

int  swb_flatten_network(swb_flat *f);          /* 0 ok, SWB_ERR_UNSUPP if element unsupported */
/* external + dry-weather inflows of the open project (Node[].extInflow / dwfInflow, Pattern[], Tseries[]);
 * blocks are owned by `f` like the network's.  Member scale / shift are left NULL. */
int  swb_flatten_inflows(swb_flat *f, swb_inflow_desc *out);
/* control rules (controls.c keeps them in file-private structs: the layouts are restated in flatten.c),
 * pump start-up / shut-off depths, orifice opening rates.  SWB_ERR_UNSUPP for rules that use named
 * variables, math expressions or rain gages. */
int  swb_flatten_controls(swb_flat *f, swb_controls_desc *out);
void swb_flat_free(swb_flat *f);
/* copy one dynamic field of the live engine (SWB_NODE_* / SWB_LINK_* ids) to/from buf */
int  swb_engine_get_field(int field, double *buf);
int  swb_engine_set_field(int field, const double *buf);
int  swb_field_len(int field, int n_nodes, int n_links, int n_pollut);
#endif
