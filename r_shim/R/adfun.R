# MakeADFun-shaped object on top of the CUDA engine: drop-in for the two TMB::MakeADFun calls of
# SDE$setup() (R/sde.R:656-669).  `map` and `random` are handled here, as TMB's R layer does.
# The Python mirror of this file (smoothsde_b200/adfun.py) is what the automated tests exercise.
MakeADFun_b200 <- function(data, parameters, map = list(), random = NULL, device = 0L, ...) {
    ptr <- .Call("ssde_make", data, as.integer(device))
    lay <- .Call("ssde_layout", ptr)                  # offsets[5], sizes[5]
    pnames <- c("log_sigma_obs", "coeff_fe", "log_lambda", "log_decay", "coeff_re")
    # log_decay is in the list for BM/OU (R/sde.R:504-507) but only enters the vector with decay terms
    full <- unlist(lapply(seq_along(pnames), function(k) if(lay[5 + k] > 0) parameters[[pnames[k]]] else NULL))
    names(full) <- rep(pnames, lay[6:10])
    # map: factor(NA) fixes an entry, equal levels tie entries together (TMB semantics)
    group <- seq_along(full)
    for(nm in names(map)) {
        idx <- which(names(full) == nm)
        f <- map[[nm]]
        group[idx] <- ifelse(is.na(f), NA, idx[match(f, f)])
    }
    is_random <- if(is.null(random)) rep(FALSE, length(full)) else names(full) %in% random
    lap <- if(any(is_random)) .Call("ssde_laplace_new", ptr) else NULL
    free <- sort(unique(group[!is.na(group)]))
    active <- free[!is_random[free]]                  # what optim sees (obj$par); random effects are integrated out
    env <- new.env()
    env$last.par <- full; env$last.par.best <- full; env$value.best <- Inf
    env$active <- active; env$random <- which(is_random & !is.na(group))   # positions in the full vector (TMB: env$random)
    scatter <- function(x) {
        p <- env$last.par                             # keeps coeff_re of the last inner optimum (warm start)
        ok <- !is.na(group) & !is_random
        p[ok] <- x[match(group[ok], active)]
        p
    }
    eval_full <- function(x, order) {
        p <- scatter(x)
        if(is.null(lap)) {
            out <- .Call("ssde_fn_gr", ptr, p, as.integer(order))
        } else {                                      # Laplace marginal: inner Newton + Cholesky on the device
            out <- .Call("ssde_laplace_fn_gr", lap, p, as.integer(order))
            p <- out$par
        }
        if(is.finite(out$value)) env$last.par <- p     # a diverged inner optimum must not become the next warm start
        if(is.finite(out$value) && out$value < env$value.best) { env$value.best <- out$value; env$last.par.best <- p }
        out
    }
    list(par = full[active],
         fn = function(x = full[active], ...) eval_full(x, 0L)$value,
         gr = function(x = full[active], ...) {
             g <- eval_full(x, 1L)$gradient
             ok <- !is.na(group) & !is_random
             matrix(tapply(g[ok], match(group[ok], active), sum), nrow = 1)
         },
         he = function(x = full[active], ...) {       # joint object only (R/sde.R:1363)
             H <- .Call("ssde_he", ptr, scatter(x))
             ok <- which(!is.na(group)); k <- match(group[ok], active)
             A <- matrix(0, length(active), length(ok)); A[cbind(k, seq_along(ok))] <- 1
             A %*% H[ok, ok] %*% t(A)
         },
         report = function(...) list(aest_all = .Call("ssde_aest", ptr, nrow(data$obs), ncol(data$a0))),
         env = env, ptr = ptr)
}

# What SDE$fit() takes from TMB::sdreport(obj, getJointPrecision = TRUE) (R/sde.R:702-719, :871-887,
# :1323, :1360-1366): par.fixed, par.random, cov.fixed and jointPrecision (rows / columns named by
# parameter, in the order of the full free parameter vector).  Same formulas as the Python mirror
# (smoothsde_b200/adfun.py: ADFun.sdreport, checked in tests/test_gpu_laplace.py):
#   no random effects: jointPrecision = exact joint Hessian (obj$he);
#   random effects:    H_f = Hessian of the Laplace marginal (central differences of its gradient,
#                      like optimHess), H_bb and G = H_{b,theta} from the exact joint Hessian at
#                      (theta, b_hat):  Q = [[H_f + G' H_bb^-1 G, G'], [G, H_bb]]   (TMB's formula).
# `obj` must come from MakeADFun_b200; `obj_joint` is the companion object without `random`
# (SDE$setup() builds both, R/sde.R:656-669).
sdreport_b200 <- function(obj, obj_joint = NULL, par.fixed = obj$env$last.par.best[obj$env$active], step = 1e-4) {
    x <- par.fixed
    nms <- names(obj$par)
    if(length(obj$env$random) == 0) {
        H <- obj$he(x)
        dimnames(H) <- list(nms, nms)
        return(structure(list(par.fixed = x, par.random = numeric(0), cov.fixed = solve(H),
                              jointPrecision = H, value = obj$fn(x)), class = "sdreport_b200"))
    }
    if(is.null(obj_joint)) stop("sdreport_b200: the joint object (random = NULL) is needed for the joint precision")
    nt <- length(x)
    Hf <- matrix(0, nt, nt)
    for(j in seq_len(nt)) {
        h <- step * max(1, abs(x[j]))
        e <- replace(numeric(nt), j, h)
        Hf[, j] <- (as.numeric(obj$gr(x + e)) - as.numeric(obj$gr(x - e))) / (2 * h)
    }
    Hf <- (Hf + t(Hf)) / 2
    f <- obj$fn(x)                                    # leaves (theta, b_hat) in env$last.par
    p <- obj$env$last.par
    free <- obj_joint$env$active                      # every free entry of the full vector, in vector order
    H <- obj_joint$he(p[free])
    is_rand <- free %in% obj$env$random
    Hbb <- H[is_rand, is_rand, drop = FALSE]
    G <- H[is_rand, !is_rand, drop = FALSE]
    Q <- H
    Q[!is_rand, !is_rand] <- Hf + t(G) %*% solve(Hbb, G)
    dimnames(Q) <- list(names(p)[free], names(p)[free])
    structure(list(par.fixed = x, par.random = p[obj$env$random], cov.fixed = solve(Hf),
                   jointPrecision = Q, value = f), class = "sdreport_b200")
}

# as.list(rep, "Estimate") (R/sde.R:707): estimates split by parameter name
as.list.sdreport_b200 <- function(x, what = "Estimate", ...) {
    est <- c(x$par.fixed, x$par.random)
    split(unname(est), factor(names(est), levels = unique(names(est))))
}
