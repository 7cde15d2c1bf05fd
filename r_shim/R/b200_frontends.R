# R front-ends over the B200 engine: same exported names and arguments as bayesSSM
# (R/bootstrap_filter.R:129, R/auxiliary_filter.R:163, R/resample_move_filter.R:190, R/pmmh.R:243).
# init_fn / transition_fn / log_likelihood_fn (/ aux_log_likelihood_fn / move_fn) take a device-model
# slot created by b200_model() (built-in models) or b200_cuda_model() (a CUDA snippet compiled by NVRTC);
# log_priors entries are device prior specs (b200_prior_*).  Plain R closures are rejected: no CPU fallback.
# Not executable in the build image (no R there); kept thin so that the behaviour lives in the C ABI, which
# is tested through the Python mirror of these functions (bayesssm_b200/filters.py, pmmh.py, sharding.py).

.b200_models <- c(nonlinear_ar = 0L, linear_gaussian = 1L, random_walk_drift = 2L,
                  sir_chain_binomial = 3L, nonlinear_ar_cos_obs = 4L, random_walk_2d = 5L,
                  sir_gillespie = 6L)
.b200_params <- list(nonlinear_ar = c("phi", "sigma_x", "sigma_y"), linear_gaussian = c("phi", "sigma_x", "sigma_y"),
                     random_walk_drift = c("mu", "sigma"), sir_chain_binomial = c("lambda", "gamma"),
                     nonlinear_ar_cos_obs = c("phi", "sigma_x", "sigma_y"), random_walk_2d = c("phi"),
                     sir_gillespie = c("lambda", "gamma"))
.b200_consts <- list(sir_chain_binomial = c("pop", "I0"), sir_gillespie = c("pop", "I0"))

.b200_slots <- function(desc) {
  slot <- function(s) structure(list(model = desc, slot = s), class = "b200_device_fn")
  list(init_fn = slot("init_fn"), transition_fn = slot("transition_fn"),
       log_likelihood_fn = slot("log_likelihood_fn"), aux_log_likelihood_fn = slot("aux_log_likelihood_fn"),
       move_fn = slot("move_fn"))
}

# built-in device models (bayesssm_b200/csrc/bssm_models.cuh)
b200_model <- function(name) {
  stopifnot(name %in% names(.b200_models))
  consts <- if (is.null(.b200_consts[[name]])) character() else .b200_consts[[name]]
  .b200_slots(list(name = name, id = .b200_models[[name]], params = .b200_params[[name]], consts = consts))
}

# user model: a CUDA snippet defining `struct UserModel` (contract at the top of bssm_models.cuh), compiled by NVRTC
# for sm_100a -- the counterpart of passing R closures (precedent: the cppFunction transition of
# vignettes/articles/detailed-overview.Rmd:408-466)
b200_cuda_model <- function(name, source, param_names, const_names = character()) {
  r <- .Call("_bayesSSM_b200_model_compile", source)
  if (r$ntheta != length(param_names) || r$nconst != length(const_names))
    stop("UserModel declares ", r$ntheta, " parameters and ", r$nconst, " constants")
  .b200_slots(list(name = name, id = r$model, params = param_names, consts = const_names))
}

# device prior specs (kind, a, b): include/bayesssm_b200.h BSSM_PRIOR_*
b200_prior_flat <- function() c(0, 0, 0)
b200_prior_normal <- function(mean = 0, sd = 1) c(1, mean, sd)
b200_prior_exponential <- function(rate = 1) c(2, rate, 0)
b200_prior_uniform <- function(min = 0, max = 1) c(3, min, max)
b200_prior_halfnormal <- function(sigma = 1) c(4, sigma, 0)

.b200_resolve <- function(...) {
  fns <- Filter(Negate(is.null), list(...))
  if (!all(vapply(fns, inherits, logical(1), "b200_device_fn")))
    stop("init_fn / transition_fn / log_likelihood_fn must be device-model slots (b200_model(), b200_cuda_model()); R closures cannot run on the GPU")
  ids <- unique(vapply(fns, function(f) f$model$id, integer(1)))
  if (length(ids) != 1) stop("operator slots belong to different device models")
  fns[[1]]$model
}

.b200_theta <- function(model, dots) {
  names_all <- c(model$params, model$consts)
  missing <- setdiff(names_all, names(dots))
  if (length(missing)) stop("missing model parameter(s): ", paste(missing, collapse = ", "))
  as.numeric(unlist(dots[names_all]))
}

.b200_filter <- function(algorithm, y, num_particles, model, obs_times, resample_algorithm, resample_fn,
                         threshold, return_particles, ..., precision = "f64", engine = "auto", carry_weights = FALSE) {
  # carry_weights = TRUE (an extension, not the reference's rule): weights carried over steps that do not resample
  checkmate::assert_count(num_particles, positive = TRUE)
  checkmate::assert_numeric(y, any.missing = FALSE)
  if (is.vector(y)) y <- matrix(y, ncol = 1)
  storage.mode(y) <- "double"   # integer observations (rpois ...) are numeric to the reference; the shim reads REAL()
  if (!is.null(obs_times)) checkmate::assert_integerish(obs_times, len = nrow(y), lower = 1, sorted = TRUE)
  theta <- .b200_theta(model, list(...))
  cfg <- list(model = model$id, algorithm = match(algorithm, c("BPF", "APF", "RMPF")) - 1L,
              resample_algorithm = match(resample_algorithm, c("SIS", "SISR", "SISAR")) - 1L,
              resample_fn = match(resample_fn, c("stratified", "systematic", "multinomial")) - 1L,
              threshold = if (is.null(threshold)) -1 else threshold, num_particles = as.integer(num_particles),
              precision = match(precision, c("f32", "f64")) - 1L,
              engine = match(engine, c("auto", "general", "persistent", "stream")) - 1L,
              seed = sample.int(.Machine$integer.max, 1), return_particles = as.integer(return_particles),
              carry_weights = as.integer(isTRUE(carry_weights)),
              obs_times = if (is.null(obs_times)) NULL else as.integer(obs_times))
  r <- .Call("_bayesSSM_b200_filter", cfg, y, theta)
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

# pmmh (R/pmmh.R:243-630): same arguments; the pilot chain, the pilot run, the tuned main chains, the transforms and
# the accept / reject step run on the device for all chains at once (bssm_pmmh_run); burn-in removal, ess() / rhat(),
# the data frame and the warnings stay here, as in the reference (R/pmmh.R:540-629).
pmmh <- function(pf_wrapper, y, m, init_fn, transition_fn, log_likelihood_fn, log_priors, pilot_init_params, burn_in,
                 num_chains = 4, obs_times = NULL, resample_algorithm = c("SISAR", "SISR", "SIS"),
                 resample_fn = c("stratified", "systematic", "multinomial"), param_transform = NULL,
                 tune_control = default_tune_control(), verbose = FALSE, return_latent_state_est = FALSE,
                 seed = NULL, num_cores = 1, ..., aux_log_likelihood_fn = NULL, move_fn = NULL, num_particles = NULL) {
  checkmate::assert_count(m, positive = TRUE)
  checkmate::assert_int(burn_in, lower = 0, upper = m - 1)
  checkmate::assert_count(num_chains, positive = TRUE)
  checkmate::assert_list(pilot_init_params, len = num_chains)
  match.arg(resample_algorithm); match.arg(resample_fn)      # validated, then unused: quirk A10 of the reference
  algorithm <- if (identical(pf_wrapper, bootstrap_filter)) 0L else if (identical(pf_wrapper, auxiliary_filter)) 1L
               else if (identical(pf_wrapper, resample_move_filter)) 2L
               else stop("pf_wrapper must be bootstrap_filter, auxiliary_filter or resample_move_filter")
  model <- .b200_resolve(init_fn, transition_fn, log_likelihood_fn, aux_log_likelihood_fn, move_fn)
  if (!setequal(names(log_priors), model$params) || !setequal(names(pilot_init_params[[1]]), model$params))
    stop("Parameters in functions do not match the names in pilot_init_params and log_priors")
  if (is.null(param_transform)) param_transform <- as.list(setNames(rep("identity", length(model$params)), model$params))
  priors <- do.call(rbind, log_priors[model$params])            # rows (kind, a, b)
  dots <- list(...)
  consts <- if (length(model$consts)) as.numeric(unlist(dots[model$consts])) else NULL
  init <- t(vapply(pilot_init_params, function(p) as.numeric(unlist(p[model$params])), numeric(length(model$params))))
  if (length(model$params) == 1) init <- matrix(init, ncol = 1)
  if (is.vector(y)) y <- matrix(y, ncol = 1)
  storage.mode(y) <- "double"   # integer observations (rpois ...) are numeric to the reference; the shim reads REAL()
  if (is.null(seed)) seed <- sample.int(.Machine$integer.max, 1)
  cfg <- list(model = model$id, algorithm = algorithm, prior_kind = as.integer(priors[, 1]), prior_a = priors[, 2],
              prior_b = priors[, 3],
              transform = as.integer(match(unlist(param_transform[model$params]), c("identity", "log", "logit")) - 1L),
              pilot_proposal_sd = rep_len(as.numeric(tune_control$pilot_proposal_sd), length(model$params)),
              pilot_n = tune_control$pilot_n, pilot_m = tune_control$pilot_m, pilot_reps = tune_control$pilot_reps,
              pilot_resample_algorithm = match(tune_control$pilot_resample_algorithm, c("SIS", "SISR", "SISAR")) - 1L,
              pilot_resample_fn = match(tune_control$pilot_resample_fn, c("stratified", "systematic", "multinomial")) - 1L,
              m = as.integer(m), num_particles = if (is.null(num_particles)) 0L else as.integer(num_particles),
              obs_times = if (is.null(obs_times)) NULL else as.integer(obs_times), consts = consts, precision = 1L,
              seed = seed, return_latent_state_est = as.integer(return_latent_state_est))
  r <- .Call("_bayesSSM_b200_pmmh", cfg, y, init)
  if (any(r$status == 5L)) stop("Initial parameter values are invalid: the log-prior is not finite (modify pilot_init_params)")
  if (any(r$status != 0L)) stop("PMMH chain failed with engine status ", paste(r$status, collapse = " "))
  keep <- (burn_in + 1):m                                          # R/pmmh.R:540-545
  chains <- lapply(seq_len(num_chains), function(c) {
    df <- as.data.frame(matrix(r$theta_chain[keep, , c], ncol = length(model$params)))
    names(df) <- model$params
    df
  })
  theta_chain <- dplyr::bind_rows(chains, .id = "chain")
  param_ess <- list(); param_rhat <- list()
  for (j in seq_along(model$params)) {                            # R/pmmh.R:570-594; ess() / rhat() below run on the device
    mat <- matrix(r$theta_chain[keep, j, ], ncol = num_chains)
    param_ess[[model$params[j]]] <- if (num_chains > 1) ess(mat) else NA
    param_rhat[[model$params[j]]] <- rhat(mat)
  }
  if (num_chains == 1) message("ESS cannot be computed with only one chain Run at least 2 chains.")
  result <- list(theta_chain = theta_chain, diagnostics = list(ess = param_ess, rhat = param_rhat))
  if (return_latent_state_est)                                     # array [T+1, d, m, chain] -> list of lists as R/pmmh.R:547-552
    result$latent_state_chain <- lapply(seq_len(num_chains), function(c) lapply(keep, function(i) drop(r$latent_state_chain[, , i, c])))
  class(result) <- "pmmh_output"
  print(result)
  if (any(unlist(param_ess) < 400, na.rm = TRUE))
    warning("Some ESS values are below 400, indicating poor mixing. Consider running the chains for more iterations.")
  if (any(unlist(param_rhat) > 1.01, na.rm = TRUE))
    warning("Some Rhat values are above 1.01, indicating that the chains have not converged.")
  result
}

# name of the CUDA device the engine runs on (errors if there is none: the engine has no CPU fallback)
b200_device_info <- function() .Call("_bayesSSM_b200_device_info")

# ---- ess() / rhat() (R/ESS.R:30-145, R/rhat.R:27-108): same inputs, messages and NA + warning behaviour; the
# variances, the autocorrelations of every chain and Geyer's truncation are computed by bssm_mcmc_diagnostics ----
.b200_diag <- function(chains, which) {
  one <- function(mat) {
    if (nrow(mat) < 2) stop("Number of iterations must be at least 2.")
    if (which == "ess" && ncol(mat) < 2) stop("Number of chains must be at least 2.")
    storage.mode(mat) <- "double"
    r <- .Call("_bayesSSM_b200_mcmc_diagnostics", mat, as.integer(which == "ess"))
    if (bitwAnd(r$flags, if (which == "ess") 1L else 2L) != 0L) {
      warning("One or more chains have zero variance.")
      return(NA)
    }
    r[[which]]
  }
  if (!is.matrix(chains) && !is.data.frame(chains))
    stop("Input must be a matrix or a data frame with a 'chain' column.")
  if (is.matrix(chains)) return(one(chains))
  if (!"chain" %in% names(chains)) stop("Data frame must contain a 'chain' column.")
  param_cols <- setdiff(names(chains), "chain")
  chain_ids <- unique(chains$chain)
  sapply(param_cols, function(param) {
    param_data <- lapply(chain_ids, function(chain) chains[[param]][chains$chain == chain])
    if (length(unique(sapply(param_data, length))) != 1) stop("Not all chains have the same number of iterations.")
    one(do.call(cbind, param_data))
  })
}
ess <- function(chains) .b200_diag(chains, "ess")
rhat <- function(chains) .b200_diag(chains, "rhat")

# ---- one filter larger than one GPU: particle-sharded over the GPUs of a box, one R process per GPU ----
# rank 0: id <- b200_shard_unique_id(); send it to the other ranks (Rmpi::mpi.bcast, a file, a socket);
# every rank: b200_shard_init(rank, world, id); then the collective b200_sharded_bootstrap_filter(...)
b200_shard_unique_id <- function() .Call("_bayesSSM_b200_shard_unique_id")
b200_shard_init <- function(rank, world, id = NULL) invisible(.Call("_bayesSSM_b200_shard_init", as.integer(rank), as.integer(world), id))
# optional, after b200_shard_init on every rank: h <- b200_shard_peer_export(); gather the raw(64) of all ranks in rank order
# (Rmpi::mpi.allgather, files); b200_shard_peer_attach(do.call(c, handles)) -- the exchange then runs inside the filter kernel
b200_shard_peer_export <- function() .Call("_bayesSSM_b200_shard_peer_export")
b200_shard_peer_attach <- function(handles) invisible(.Call("_bayesSSM_b200_shard_peer_attach", handles))
b200_sharded_bootstrap_filter <- function(y, num_particles, init_fn, transition_fn, log_likelihood_fn,
                                          resample_algorithm = c("SISAR", "SISR", "SIS"),
                                          resample_fn = c("stratified", "systematic"), threshold = NULL,
                                          capacity_factor = 1.5, seed = 1L, ...) {
  model <- .b200_resolve(init_fn, transition_fn, log_likelihood_fn)
  resample_algorithm <- match.arg(resample_algorithm); resample_fn <- match.arg(resample_fn)
  if (is.vector(y)) y <- matrix(y, ncol = 1)
  storage.mode(y) <- "double"   # integer observations (rpois ...) are numeric to the reference; the shim reads REAL()
  cfg <- list(model = model$id, resample_algorithm = match(resample_algorithm, c("SIS", "SISR", "SISAR")) - 1L,
              resample_fn = match(resample_fn, c("stratified", "systematic")) - 1L,
              threshold = if (is.null(threshold)) -1 else threshold, num_particles = as.integer(num_particles),
              precision = 0L, seed = seed, capacity_factor = capacity_factor)   # the same seed on every rank
  r <- .Call("_bayesSSM_b200_shard_filter", cfg, y, .b200_theta(model, list(...)))
  out <- list(state_est = r$state_est, ess = r$ess, loglike = r$loglike, loglike_history = r$loglike_history, algorithm = "BPF")
  if (!r$early_exit) out$resample_algorithm <- resample_algorithm
  out
}
