/* Constants of Ceres Solver 1.9.0 that the reference reaches through its third-party dependency
 * (pinned only by /root/reference/.travis.yml:33 and README.md:3; the sources are NOT in the reference
 * tree and not in the build container).  Every value below is RECALLED from Ceres 1.9.0's published
 * sources -- file and symbol named per line -- and is the one place to correct when a Ceres checkout is
 * available.  Used by the product (rsba_cuda_default_options, reproj_math.cuh, lm_solver.cu), by the
 * oracle's Ceres stand-in (oracle/shim/ceres/rotation.h) and mirrored in oracle/lm_oracle.py (Options).
 * Plain C: includable from C, C++, CUDA. */
#ifndef RSBA_CERES_CONSTANTS_H_
#define RSBA_CERES_CONSTANTS_H_

/* ceres/rotation.h, AngleAxisRotatePoint: Rodrigues' formula when theta^2 > epsilon, else the first-order
 * form  p + r x p.  (Releases differ between `> 0.0` and `> std::numeric_limits<double>::epsilon()`; for
 * theta^2 in (0, eps] the two branches agree to < 1e-16 in value and 1e-8 in the derivative.) */
#define RSBA_ANGLE_AXIS_EPS 2.220446049250313e-16

/* ceres/solver.h, Solver::Options defaults (trust-region minimizer, LEVENBERG_MARQUARDT) */
#define RSBA_CERES_MAX_NUM_ITERATIONS 50              /* also what CeresHandler.h:405 sets */
#define RSBA_CERES_INITIAL_TRUST_REGION_RADIUS 1e4
#define RSBA_CERES_MAX_TRUST_REGION_RADIUS 1e16
#define RSBA_CERES_MIN_TRUST_REGION_RADIUS 1e-32
#define RSBA_CERES_MIN_RELATIVE_DECREASE 1e-3
#define RSBA_CERES_MIN_LM_DIAGONAL 1e-6
#define RSBA_CERES_MAX_LM_DIAGONAL 1e32
#define RSBA_CERES_FUNCTION_TOLERANCE 1e-6
#define RSBA_CERES_GRADIENT_TOLERANCE 1e-10
#define RSBA_CERES_PARAMETER_TOLERANCE 1e-8
#define RSBA_CERES_MAX_NUM_CONSECUTIVE_INVALID_STEPS 5
#define RSBA_CERES_JACOBI_SCALING 1

/* ceres/levenberg_marquardt_strategy.cc: StepAccepted  radius /= max(1/3, 1 - (2 rho - 1)^3), decrease
 * factor back to 2;  StepRejected / StepIsInvalid  radius /= decrease_factor, decrease_factor *= 2 */
#define RSBA_CERES_LM_MIN_RADIUS_SHRINK (1.0 / 3.0)
#define RSBA_CERES_LM_INITIAL_DECREASE_FACTOR 2.0

/* ceres/trust_region_minimizer.cc: Jacobi scaling 1 / (1 + sqrt(column norm^2)), fixed at iteration 0;
 * a step is valid iff model_cost_change > 0; a linear-solver failure or an invalid step counts towards
 * max_num_consecutive_invalid_steps and shrinks the radius like a rejected step. */

/* ceres/solver.h, line search defaults (used by the trust-region minimizer on bound-constrained problems):
 * ARMIJO, sufficient_function_decrease 1e-4, the search starts at step size 1 */
#define RSBA_CERES_ARMIJO_SUFFICIENT_DECREASE 1e-4

/* ceres/loss_function.h, HuberLoss(a): rho(s) = s for s <= a^2, else 2 a sqrt(s) - a^2 */

#endif  /* RSBA_CERES_CONSTANTS_H_ */
