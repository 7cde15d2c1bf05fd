# R front-ends over the B200 engine: same exported names and arguments as bayesSSM
# (R/bootstrap_filter.R:129, R/auxiliary_filter.R:163, R/resample_move_filter.R:190, R/pmmh.R:243).
# init_fn / transition_fn / log_likelihood_fn (/ aux_log_likelihood_fn / move_fn) take a device-model
# slot created by b200_model(); plain R closures are rejected (no CPU fallback).
# Not executable in the build image (no R there); kept logic-free so that the behaviour lives in the
# C ABI, which is tested through the Python mirror of these functions (bayesssm_b200/filters.py, pmmh.py).

.b200_models <- c(nonlinear_ar = 0L, linear_gaussian = 1L, random_walk_drift = 2L,
                  sir_chain_binomial = 3L, nonlinear_ar_cos_obs = 4L, random_walk_2d = 5L)
.b200_params <- list(nonlinear_ar = c("phi", "sigma_x", "sigma_y"), linear_gaussian = c("phi", "sigma_x", "sigma_y"),
                     random_walk_drift = c("mu", "sigma"), sir_chain_binomial = c("lambda", "gamma", "pop", "I0"),
                     nonlinear_ar_cos_obs = c("phi", "sigma_x", "sigma_y"), random_walk_2d = c("phi"))

b200_model <- function(name) {
  stopifnot(name %in% names(.b200_models))
  slot <- function(s) structure(list(model = name, slot = s), class = "b200_device_fn")
  list(init_fn = slot("init_fn"), transition_fn = slot("transition_fn"),
       log_likelihood_fn = slot("log_likelihood_fn"), aux_log_likelihood_fn = slot("aux_log_likelihood_fn"),
       move_fn = slot("move_fn"))
}

.b200_resolve <- function(...) {
  fns <- Filter(Negate(is.null), list(...))
  if (!all(vapply(fns, inherits, logical(1), "b200_device_fn")))
    stop("init_fn / transition_fn / log_likelihood_fn must be device-model slots (b200_model()); R closures cannot run on the GPU")
  m <- unique(vapply(fns, function(f) f$model, character(1)))
  if (length(m) != 1) stop("operator slots belong to different device models")
  m
}

.b200_filter <- function(algorithm, y, num_particles, model, obs_times, resample_algorithm, resample_fn,
                         threshold, return_particles, ...) {
  checkmate::assert_count(num_particles, positive = TRUE)
  checkmate::assert_numeric(y, any.missing = FALSE)
  if (is.vector(y)) y <- matrix(y, ncol = 1)
  if (!is.null(obs_times)) checkmate::assert_integerish(obs_times, len = nrow(y), lower = 1, sorted = TRUE)
  dots <- list(...)
  theta <- unlist(dots[.b200_params[[model]]])
  if (length(theta) != length(.b200_params[[model]])) stop("missing model parameter(s)")
  cfg <- list(model = .b200_models[[model]], algorithm = match(algorithm, c("BPF", "APF", "RMPF")) - 1L,
              resample_algorithm = match(resample_algorithm, c("SIS", "SISR", "SISAR")) - 1L,
              resample_fn = match(resample_fn, c("stratified", "systematic", "multinomial")) - 1L,
              threshold = if (is.null(threshold)) -1 else threshold, num_particles = as.integer(num_particles),
              precision = 1L, seed = sample.int(.Machine$integer.max, 1), return_particles = as.integer(return_particles),
              obs_times = if (is.null(obs_times)) NULL else as.integer(obs_times))
  r <- .Call("_bayesSSM_b200_filter", cfg, y, as.numeric(theta))
  out <- list(state_est = r$state_est, ess = r$ess, loglike = r$loglike, loglike_history = r$loglike_history,
              algorithm = algorithm)
  if (!r$early_exit) out$resample_algorithm <- resample_algorithm
  if (return_particles) { out$particles_history <- r$particles_history; out$weights_history <- r$weights_history }
  out
}

bootstrap_filter <- function(y, num_particles, init_fn, transition_fn, log_likelihood_fn, obs_times = NULL,
                             resample_algorithm = c("SISAR", "SISR", "SIS"),
                             resample_fn = c("stratified", "systematic", "multinomial"),
                             threshold = NULL, return_particles = TRUE, ...) {
  .b200_filter("BPF", y, num_particles, .b200_resolve(init_fn, transition_fn, log_likelihood_fn), obs_times,
               match.arg(resample_algorithm), match.arg(resample_fn), threshold, return_particles, ...)
}

auxiliary_filter <- function(y, num_particles, init_fn, transition_fn, log_likelihood_fn, aux_log_likelihood_fn,
                             obs_times = NULL, resample_algorithm = c("SISAR", "SISR", "SIS"),
                             resample_fn = c("stratified", "systematic", "multinomial"),
                             threshold = NULL, return_particles = TRUE, ...) {
  .b200_filter("APF", y, num_particles, .b200_resolve(init_fn, transition_fn, log_likelihood_fn, aux_log_likelihood_fn),
               obs_times, match.arg(resample_algorithm), match.arg(resample_fn), threshold, return_particles, ...)
}

resample_move_filter <- function(y, num_particles, init_fn, transition_fn, log_likelihood_fn, move_fn, obs_times = NULL,
                                 resample_fn = c("stratified", "systematic", "multinomial"),
                                 threshold = NULL, return_particles = TRUE, ...) {
  dots <- list(...); dots$resample_algorithm <- NULL   # R/resample_move_filter.R:213-216
  do.call(.b200_filter, c(list("RMPF", y, num_particles, .b200_resolve(init_fn, transition_fn, log_likelihood_fn, move_fn),
                               obs_times, "SISR", match.arg(resample_fn), threshold, return_particles), dots))
}
