"""pmmh() and default_tune_control() with the reference's arguments and return object
(R/pmmh.R:33-58,243-630; pilot R/pmmh_tuning.R), on the device-resident PMMH engine.

log_priors: the reference takes R closures; here each prior is a device prior spec from
`priors` below (normal / exponential / uniform / half_normal / flat), evaluated on the GPU."""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _native as nat
from . import filters as _filters
from .diagnostics import _ZERO_VAR, device_diagnostics
from .models import resolve_model


class priors:
    """Device log-prior specs (Appendix D of SURVEY.md lists the priors the reference uses)."""
    @staticmethod
    def flat(): return (nat.PRIOR_FLAT, 0.0, 0.0)
    @staticmethod
    def normal(mean=0.0, sd=1.0): return (nat.PRIOR_NORMAL, float(mean), float(sd))
    @staticmethod
    def exponential(rate=1.0): return (nat.PRIOR_EXP, float(rate), 0.0)
    @staticmethod
    def uniform(lo=0.0, hi=1.0): return (nat.PRIOR_UNIF, float(lo), float(hi))
    @staticmethod
    def half_normal(sigma=1.0): return (nat.PRIOR_HALFNORMAL, float(sigma), 0.0)


def default_tune_control(pilot_proposal_sd=0.5, pilot_n=100, pilot_m=2000, pilot_target_var=1, pilot_burn_in=500,
                         pilot_reps=100, pilot_resample_algorithm=("SISAR", "SISR", "SIS"),
                         pilot_resample_fn=("stratified", "systematic", "multinomial")):
    """R/pmmh.R:33-58 (pilot_target_var and pilot_burn_in are validated but never read, quirk A11)."""
    def number(v, name):
        if not (isinstance(v, (int, float)) and np.isfinite(v) and v >= 0):
            raise ValueError(f"Assertion on '{name}' failed: Must be a finite number >= 0")
    def count(v, name):
        if not (isinstance(v, (int, np.integer)) and v > 0):
            raise ValueError(f"Assertion on '{name}' failed: Must be a positive count")
    number(pilot_proposal_sd, "pilot_proposal_sd"); count(pilot_n, "pilot_n"); count(pilot_m, "pilot_m")
    number(pilot_target_var, "pilot_target_var"); count(pilot_burn_in, "pilot_burn_in"); count(pilot_reps, "pilot_reps")
    return {"pilot_proposal_sd": pilot_proposal_sd, "pilot_n": pilot_n, "pilot_m": pilot_m,
            "pilot_target_var": pilot_target_var, "pilot_burn_in": pilot_burn_in, "pilot_reps": pilot_reps,
            "pilot_resample_algorithm": _filters._match_arg(pilot_resample_algorithm, ("SISAR", "SISR", "SIS"), "pilot_resample_algorithm"),
            "pilot_resample_fn": _filters._match_arg(pilot_resample_fn, ("stratified", "systematic", "multinomial"), "pilot_resample_fn")}


_WRAPPERS = {_filters.bootstrap_filter: nat.BPF, _filters.auxiliary_filter: nat.APF,
             _filters.resample_move_filter: nat.RMPF, _filters.particle_filter: nat.BPF}
_TRANSFORMS = {"identity": nat.TR_IDENTITY, "log": nat.TR_LOG, "logit": nat.TR_LOGIT}


class PmmhOutput(dict):
    """`pmmh_output` (R/pmmh.R:599-608): theta_chain (DataFrame with a leading character `chain` column),
    diagnostics = {ess, rhat}, optional latent_state_chain.  str() is print.pmmh_output (R/print.R:30-66),
    summary() is summary.pmmh_output (R/summary.R:28-54)."""

    def _param_names(self):
        return [c for c in self["theta_chain"].columns if c != "chain"]

    def summary(self):
        """Data frame indexed by parameter: mean, sd, median, 2.5%, 97.5% (stats::quantile's default type 7 = numpy's
        linear interpolation), ESS, Rhat -- unrounded (R/summary.R:33-53)."""
        import pandas as pd
        tc = self["theta_chain"]
        rows = {}
        for name in self._param_names():
            x = tc[name].to_numpy(dtype=np.float64)
            lo, hi = np.quantile(x, [0.025, 0.975])
            rows[name] = {"mean": x.mean(), "sd": x.std(ddof=1), "median": np.median(x), "2.5%": lo, "97.5%": hi,
                          "ESS": self["diagnostics"]["ess"][name], "Rhat": self["diagnostics"]["rhat"][name]}
        return pd.DataFrame.from_dict(rows, orient="index")

    def __str__(self):
        """"PMMH Results Summary:" and one row per parameter; statistics rounded to 2 digits, ESS floored, Rhat
        rounded to 3 (R/print.R:34-63)."""
        import pandas as pd
        sm = self.summary()
        tab = pd.DataFrame({"Parameter": sm.index, "Mean": sm["mean"].round(2).to_numpy(), "SD": sm["sd"].round(2).to_numpy(),
                            "Median": sm["median"].round(2).to_numpy(), "2.5%": sm["2.5%"].round(2).to_numpy(),
                            "97.5%": sm["97.5%"].round(2).to_numpy(), "ESS": np.floor(sm["ESS"].to_numpy(dtype=np.float64)),
                            "Rhat": sm["Rhat"].round(3).to_numpy()})
        if np.isfinite(tab["ESS"]).all():
            tab["ESS"] = tab["ESS"].astype(np.int64)
        return "PMMH Results Summary:\n" + tab.to_string(index=False)


def run_chains(ctx, model, algorithm, y, init_theta, prior_specs, transforms, tune_control, m, seed,
               chain_id_base=0, fixed_num_particles=0, consts=(), obs_times=None, precision=nat.F64,
               skip_pilot=False, proposal_chol=None, engine=nat.ENGINE_AUTO, return_latent_state_est=False):
    """Raw engine call for this process's shard of chains (used by pmmh() and by the multi-GPU bench)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    T, dy = y.shape
    init_theta = np.ascontiguousarray(init_theta, dtype=np.float64)
    Cn, p = init_theta.shape
    cfg = nat.PmmhConfig()
    pk = np.ascontiguousarray([s[0] for s in prior_specs], dtype=np.int32)
    pa = np.ascontiguousarray([s[1] for s in prior_specs], dtype=np.float64)
    pb = np.ascontiguousarray([s[2] for s in prior_specs], dtype=np.float64)
    tr = np.ascontiguousarray(transforms, dtype=np.int32)
    sd = np.ascontiguousarray(np.resize(np.asarray(tune_control["pilot_proposal_sd"], dtype=np.float64), p))
    cfg.model, cfg.algorithm, cfg.p = model.model_id, algorithm, p
    cfg.prior_kind, cfg.prior_a, cfg.prior_b = pk.ctypes.data_as(nat.c_int_p), pa.ctypes.data_as(nat.c_double_p), pb.ctypes.data_as(nat.c_double_p)
    cfg.transform, cfg.pilot_proposal_sd = tr.ctypes.data_as(nat.c_int_p), sd.ctypes.data_as(nat.c_double_p)
    cfg.pilot_n, cfg.pilot_m, cfg.pilot_reps = int(tune_control["pilot_n"]), int(tune_control["pilot_m"]), int(tune_control["pilot_reps"])
    cfg.pilot_resample_algorithm = nat.RESAMPLE_ALGORITHMS[tune_control["pilot_resample_algorithm"]]
    cfg.pilot_resample_fn = nat.RESAMPLE_FNS[tune_control["pilot_resample_fn"]]
    cfg.m, cfg.num_chains, cfg.chain_id_base = int(m), Cn, int(chain_id_base)
    cfg.fixed_num_particles = int(fixed_num_particles)
    cfg.num_obs, cfg.dy = T, dy
    ot = None
    if obs_times is not None:
        ot = np.ascontiguousarray(obs_times, dtype=np.int32)
        cfg.obs_times = ot.ctypes.data_as(nat.c_int_p)
    cs = np.ascontiguousarray(consts if len(consts) else [0.0], dtype=np.float64)
    cfg.consts, cfg.nconst = cs.ctypes.data_as(nat.c_double_p), len(consts)
    cfg.precision, cfg.seed = precision, int(seed)
    cfg.skip_pilot = int(skip_pilot)
    chol_in = None
    if proposal_chol is not None:
        chol_in = np.ascontiguousarray(proposal_chol, dtype=np.float64)
        cfg.proposal_chol_in = chol_in.ctypes.data_as(nat.c_double_p)
    cfg.engine = engine
    pm = 1 if skip_pilot else cfg.pilot_m
    reps = 1 if skip_pilot else cfg.pilot_reps
    out = {"pilot_theta_chain": np.zeros((Cn, pm, p)), "pilot_loglike_chain": np.zeros((Cn, pm)),
           "pilot_theta_mean": np.zeros((Cn, p)), "pilot_theta_cov": np.zeros((Cn, p, p)),
           "pilot_loglikes": np.zeros((Cn, reps)), "proposal_chol": np.zeros((Cn, p, p)),
           "theta_chain": np.zeros((Cn, int(m), p)), "loglike_chain": np.zeros((Cn, int(m)))}
    res = nat.PmmhResult()
    for k, v in out.items():
        setattr(res, k, v.ctypes.data_as(nat.c_double_p))
    for k in ("target_n", "n_accept", "status"):
        out[k] = np.zeros(Cn, dtype=np.int32)
        setattr(res, k, out[k].ctypes.data_as(nat.c_int32_p))
    if return_latent_state_est:
        cfg.return_latent_state_est = 1
        out["latent_state_chain"] = np.zeros((Cn, int(m), T + 1, model.dim))
        res.latent_state_chain = out["latent_state_chain"].ctypes.data_as(nat.c_double_p)
    nat.check(ctx.lib.bssm_pmmh_run(ctx.handle, C.byref(cfg), y.ctypes.data_as(nat.c_double_p),
                                    init_theta.ctypes.data_as(nat.c_double_p), C.byref(res)))
    out["pilot_ms"], out["main_ms"] = res.pilot_ms, res.main_ms
    out["main_resampled_fraction"] = res.main_resampled_fraction
    return out


def pmmh(pf_wrapper, y, m, init_fn, transition_fn, log_likelihood_fn, log_priors, pilot_init_params, burn_in,
         num_chains=4, obs_times=None, resample_algorithm=("SISAR", "SISR", "SIS"),
         resample_fn=("stratified", "systematic", "multinomial"), param_transform=None, tune_control=None,
         verbose=False, return_latent_state_est=False, seed=None, num_cores=1, *, aux_log_likelihood_fn=None,
         move_fn=None, num_particles=None, precision="f64", ctx=None, print_result=True, **params):
    """Particle Marginal Metropolis-Hastings (R/pmmh.R:243-630).

    `pf_wrapper` is bootstrap_filter / auxiliary_filter / resample_move_filter (function identity selects
    the algorithm).  `num_particles` (not in the reference) overrides the tuned target_n, which the
    reference clamps to [50, 1000] (R/pmmh_tuning.R:54-57).  `resample_algorithm` / `resample_fn` are
    validated and then unused, exactly as in the reference (quirk A10).  `num_cores` is accepted and
    ignored: all chains advance together on the GPU.  Remaining **params are fixed model constants."""
    y_arr = np.asarray(y, dtype=np.float64)
    if y_arr.size == 0 or np.isnan(y_arr).any():
        raise ValueError("Assertion on 'y' failed: Must be numeric without missing values")
    if not (isinstance(m, (int, np.integer)) and m >= 1):
        raise ValueError("Assertion on 'm' failed: Must be >= 1")
    if not (isinstance(burn_in, (int, np.integer)) and 0 <= burn_in <= m - 1):
        raise ValueError(f"Assertion on 'burn_in' failed: Must be in [0, {m - 1}]")
    if not (isinstance(num_chains, (int, np.integer)) and num_chains >= 1):
        raise ValueError("Assertion on 'num_chains' failed: Must be >= 1")
    if not (isinstance(num_cores, (int, np.integer)) and num_cores >= 1):
        raise ValueError("Assertion on 'num_cores' failed: Must be >= 1")
    if not isinstance(pilot_init_params, (list, tuple)) or len(pilot_init_params) != num_chains:
        raise ValueError(f"Assertion on 'pilot_init_params' failed: Must have length {num_chains}")
    names0 = list(pilot_init_params[0].keys())
    if any(list(d.keys()) != names0 for d in pilot_init_params):
        raise ValueError("Assertion on 'pilot_init_params' failed: Must be TRUE")
    if len(names0) == 0:
        raise ValueError("pilot_init_params must contain at least one parameter.")
    _filters._match_arg(resample_algorithm, ("SISAR", "SISR", "SIS"), "resample_algorithm")
    _filters._match_arg(resample_fn, ("stratified", "systematic", "multinomial"), "resample_fn")
    if pf_wrapper not in _WRAPPERS:
        raise TypeError("pf_wrapper must be bootstrap_filter, auxiliary_filter or resample_move_filter")
    algorithm = _WRAPPERS[pf_wrapper]
    model = resolve_model(init_fn, transition_fn, log_likelihood_fn, aux_log_likelihood_fn, move_fn)
    # .check_params_match (R/utils.R:15-72): parameter names of the model, the priors and the inits agree
    if set(log_priors.keys()) != set(model.param_names) or set(names0) != set(model.param_names):
        raise ValueError("Parameters in functions do not match the names in pilot_init_params and log_priors: "
                         f"model takes {list(model.param_names)}")
    consts = []
    for cn in model.const_names:
        if cn not in params:
            raise ValueError(f"missing model constant '{cn}'")
        consts.append(float(params[cn]))
    if param_transform is None:
        param_transform = {n: "identity" for n in log_priors}
    elif isinstance(param_transform, dict):
        if not all(n in param_transform for n in log_priors):
            raise ValueError("param_transform must include an entry for every parameter in log_priors.")
        if any(v not in _TRANSFORMS for v in param_transform.values()):
            warnings.warn("Only 'log', 'logit', and 'identity' transformations are supported. Using 'identity' for invalid entries.")
            param_transform = {k: (v if v in _TRANSFORMS else "identity") for k, v in param_transform.items()}
    else:
        raise ValueError("param_transform must be a list.")
    tune_control = dict(tune_control or default_tune_control())
    order = list(model.param_names)
    prior_specs = [log_priors[n] for n in order]
    for s in prior_specs:
        if not (isinstance(s, tuple) and len(s) == 3):
            raise TypeError("log_priors entries must be device prior specs (bayesssm_b200.priors.*); host closures "
                            "cannot run inside the engine")
    transforms = [_TRANSFORMS[param_transform[n]] for n in order]
    init_theta = np.array([[float(d[n]) for n in order] for d in pilot_init_params], dtype=np.float64)
    if seed is None:
        seed = int(np.random.default_rng().integers(1, 2**31 - 1))
    ctx = ctx or nat.default_context()
    if verbose:
        print(f"Running {num_chains} chain(s) on the GPU: pilot {tune_control['pilot_m']} iterations at "
              f"{tune_control['pilot_n']} particles, then {m} iterations")
    prec = _filters._precision(precision)
    try:
        out = run_chains(ctx, model, algorithm, y_arr, init_theta, prior_specs, transforms, tune_control, m, seed,
                         fixed_num_particles=int(num_particles or 0), consts=consts, obs_times=obs_times, precision=prec,
                         return_latent_state_est=bool(return_latent_state_est))
    except nat.EngineError as e:
        raise RuntimeError(str(e)) from e
    if (out["status"] == nat.ERR_PRIOR_INIT).any():
        raise ValueError("Initial parameter values are invalid: the log-prior is not finite "
                         "(modify pilot_init_params)")  # R/pmmh_tuning.R:136-142
    if (out["status"] != 0).any():
        raise RuntimeError(f"PMMH chain failed with engine status {out['status'].tolist()}")
    if verbose:
        for c in range(num_chains):
            print(f"Chain {c + 1}: pilot mean {np.round(out['pilot_theta_mean'][c], 4).tolist()}, "
                  f"target_n {int(out['target_n'][c])}, acceptance {out['n_accept'][c] / max(m - 1, 1):.3f}")
    import pandas as pd
    post = out["theta_chain"][:, burn_in:, :]                  # R/pmmh.R:540-545
    frames = []
    for c in range(num_chains):
        df = pd.DataFrame(post[c], columns=order)
        df.insert(0, "chain", str(c + 1))                      # bind_rows(.id = "chain"): character ids
        frames.append(df)
    theta_chain = pd.concat(frames, ignore_index=True)
    # R/pmmh.R:570-594: ess() (two or more chains) and rhat() of every parameter -- one device call over the
    # draws as they came back, burn-in skipped there
    param_ess = {name: float("nan") for name in order}
    param_rhat = dict(param_ess)
    if m - burn_in >= 2:
        dg = device_diagnostics(out["theta_chain"], burn_in, want_ess=num_chains > 1, ctx=ctx)
        for j, name in enumerate(order):
            param_ess[name], param_rhat[name] = float(dg["ess"][j]), float(dg["rhat"][j])
        if np.any(dg["flags"] & (3 if num_chains > 1 else 2)):
            warnings.warn(_ZERO_VAR)
    if num_chains == 1:
        print("ESS cannot be computed with only one chain Run at least 2 chains.")
    result = PmmhOutput(theta_chain=theta_chain, diagnostics={"ess": param_ess, "rhat": param_rhat})
    result["acceptance_rate"] = out["n_accept"] / max(m - 1, 1)
    result["target_n"] = out["target_n"].copy()
    result["timing_ms"] = {"pilot": out["pilot_ms"], "main": out["main_ms"]}
    if return_latent_state_est:
        # R/pmmh.R:547-552,604-606: per chain a list of the post-burn-in state estimates (vectors of T+1, or
        # (T+1) x d matrices for multi-dimensional states)
        lat = out["latent_state_chain"][:, burn_in:]
        d = lat.shape[-1]
        result["latent_state_chain"] = [[(lat[c, i, :, 0].copy() if d == 1 else lat[c, i].copy()) for i in range(lat.shape[1])]
                                        for c in range(num_chains)]
    if print_result:
        print(result)
    if any(np.isfinite(v) and v < 400 for v in param_ess.values()):
        warnings.warn("Some ESS values are below 400, indicating poor mixing. Consider running the chains for more iterations.")
    if any(np.isfinite(v) and v > 1.01 for v in param_rhat.values()):
        warnings.warn("\nSome Rhat values are above 1.01, indicating that the chains have not converged. \n"
                      "Consider running the chains for more iterations and/or increase burn_in.")
    return result
