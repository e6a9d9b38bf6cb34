/*
 * swmm_b200_seam.h -- the reference-facing half of the drop-in boundary.
 *
 * libswmm5_b200_seam.so exports the seven routing-seam functions of the reference engine with the
 * reference's own names, signatures and error behaviour (ErrorCode set through
 * report_writeErrorMsg, early return when ErrorCode != 0; non-convergence is not an error).
 * Each declaration cites the reference interface it replaces.  The implementation
 * (stormwater-management-model_b200/seam/seam.c) flattens the engine's global objects and drives
 * the plain C-ABI of swmm_b200.h.  See INTEGRATION.md for the two ways to bind it.
 */
#ifndef SWMM_B200_SEAM_H
#define SWMM_B200_SEAM_H
#ifdef __cplusplus
extern "C" {
#endif

void   dynwave_validate(void);                   /* funcs.h:229  dynwave.c:177  caller project.c:261  */
void   dynwave_init(void);                       /* funcs.h:230  dynwave.c:117  caller flowrout.c:87  */
void   dynwave_close(void);                      /* funcs.h:231  dynwave.c:165  caller flowrout.c:114 */
double dynwave_getRoutingStep(double fixedStep); /* funcs.h:232  dynwave.c:195  caller flowrout.c:129 */
int    dynwave_execute(double tStep);            /* funcs.h:233  dynwave.c:224  caller flowrout.c:167;
                                                    returns the Picard iterations used               */
void   qualrout_init(void);                      /* funcs.h:236  qualrout.c:63  caller routing.c:122  */
void   qualrout_execute(double tStep);           /* funcs.h:237  qualrout.c:100 caller routing.c:248  */

#ifdef __cplusplus
}
#endif
#endif
