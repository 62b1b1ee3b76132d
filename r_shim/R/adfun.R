# MakeADFun-shaped object on top of the CUDA engine: drop-in for the two TMB::MakeADFun calls of
# SDE$setup() (R/sde.R:656-669).  `map` and `random` are handled here, as TMB's R layer does.
# The Python mirror of this file (smoothsde_b200/adfun.py) is what the automated tests exercise.
MakeADFun_b200 <- function(data, parameters, map = list(), random = NULL, device = 0L, ...) {
    ptr <- .Call("ssde_make", data, as.integer(device))
    lay <- .Call("ssde_layout", ptr)                  # offsets[4], sizes[4]
    full <- unlist(parameters[c("log_sigma_obs", "coeff_fe", "log_lambda", "coeff_re")])
    names(full) <- rep(c("log_sigma_obs", "coeff_fe", "log_lambda", "coeff_re"), lay[5:8])
    # map: factor(NA) fixes an entry, equal levels tie entries together (TMB semantics)
    group <- seq_along(full)
    for(nm in names(map)) {
        idx <- which(names(full) == nm)
        f <- map[[nm]]
        group[idx] <- ifelse(is.na(f), NA, idx[match(f, f)])
    }
    if(!is.null(random)) stop("random = 'coeff_re' needs the Laplace layer (see smoothsde_b200/laplace.py)")
    active <- sort(unique(group[!is.na(group)]))
    scatter <- function(x) { p <- full; ok <- !is.na(group); p[ok] <- x[match(group[ok], active)]; p }
    env <- new.env()
    env$last.par <- full; env$last.par.best <- full; env$value.best <- Inf
    eval_full <- function(x, order) {
        p <- scatter(x)
        out <- .Call("ssde_fn_gr", ptr, p, as.integer(order))
        env$last.par <- p
        if(is.finite(out$value) && out$value < env$value.best) { env$value.best <- out$value; env$last.par.best <- p }
        out
    }
    list(par = full[active],
         fn = function(x = full[active], ...) eval_full(x, 0L)$value,
         gr = function(x = full[active], ...) {
             g <- eval_full(x, 1L)$gradient
             ok <- !is.na(group)
             matrix(tapply(g[ok], match(group[ok], active), sum), nrow = 1)
         },
         report = function(...) list(aest_all = .Call("ssde_aest", ptr, nrow(data$obs), ncol(data$obs))),
         env = env, ptr = ptr)
}
