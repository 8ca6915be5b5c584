"""Affine-invariant stretch-move ensemble sampler -- host mirror of the reference's
``magpy_rv.mcmc`` (class MCMC :34, run_MCMC :676).

The proposal (``split_step``), accept/reject (``compare``) and bookkeeping (``reset``) logic stays
on the host and consumes the two global RNG streams (Python ``random`` and legacy ``np.random``)
in exactly the reference's call order, so a shared
``random.seed(s); np.random.seed(s)`` reproduces the reference's chains step for step.  The
per-walker likelihood loop of the reference (``MCMC.compute``, mcmc.py:410-461, and the initial
loop of ``__init__``, mcmc.py:221-261) is replaced by ONE batched GPU call per iteration
(``mrv_logl_batch``); with ``torch.distributed`` initialised the ensemble is sharded across ranks
and only log-probabilities and status words are all-gathered (``sharding.py``).
"""
import os
import random
import sys
import time

import numpy as np

from . import auxiliary as aux
from . import engine
from . import mcmc_aux as mcmcx
from . import models as modl
from . import parameters as par

get_model = mcmcx.get_model
numb_param_per_model = mcmcx.numb_param_per_model
parameter_check = mcmcx.parameter_check


def _default_backend(t, rv, rv_err, kernel_name, model_name, prior_list, hparam_names, modpar_names, flags):
    """The batched GPU evaluator for one MCMC problem (sharded when torch.distributed is up)."""
    from .sharding import ShardedEvaluator
    return ShardedEvaluator(t, rv, rv_err, kernel_name, model_name, prior_list, hparam_names, modpar_names, flags)


def _randint0(n):
    """random.randint(0, n - 1): randint -> randrange -> _randbelow(n) on the module's hidden
    instance; calling the last step directly consumes the Python RNG stream identically."""
    if n <= 0:
        return random.randint(0, n - 1)     # raises exactly as the reference's call does
    return _randbelow(n)


_randbelow = getattr(getattr(random, "_inst", None), "_randbelow", None) or (lambda n: random.randint(0, n - 1))
# Random._randbelow_with_getrandbits spelled out (k = n.bit_length(); draw k bits until < n) saves two Python
# calls per number; only when the module's hidden instance is the stock Mersenne Twister class
_STOCK_RANDOM = type(getattr(random, "_inst", None)) is random.Random


def _randints0(n, count):
    """`count` successive random.randint(0, n - 1) values as an index array."""
    if n <= 0 or not _STOCK_RANDOM:
        return np.array([_randint0(n) for _ in range(count)], dtype=np.intp)
    getrandbits, k, out = random.getrandbits, n.bit_length(), []
    for _ in range(count):
        r = getrandbits(k)
        while r >= n:
            r = getrandbits(k)
        out.append(r)
    return np.array(out, dtype=np.intp)


def _shuffle(x):
    """random.shuffle(x) for a list (Fisher-Yates from the top, one _randbelow(i + 1) per element)."""
    if not _STOCK_RANDOM or sys.version_info < (3, 9):
        random.shuffle(x)
        return
    getrandbits = random.getrandbits
    for i in range(len(x) - 1, 0, -1):
        n = i + 1
        k = n.bit_length()
        j = getrandbits(k)
        while j >= n:
            j = getrandbits(k)
        x[i], x[j] = x[j], x[i]


def _parameter_check_rows(rows, names, check_args=()):
    """Row-wise ``mcmc_aux.parameter_check`` (reference mcmc_aux.py:187-231) for a [n, nm] array."""
    ok = np.ones(rows.shape[0], dtype=bool)
    if check_args:      # Rstar and Mstar given: the reference's orbit_check path (AttributeError, F10)
        for i in range(rows.shape[0]):
            ok[i] = parameter_check(rows[i], names, *check_args)
        return ok
    o = 0
    for name in names:
        if name.startswith(('kepl', 'Kepl')):
            bad = (rows[:, o] < 0.) | (rows[:, o + 1] < 0.) | (rows[:, o + 4] < 0.)
            bad |= (rows[:, o + 2] ** 2 + rows[:, o + 3] ** 2) > 1.
            ok &= ~bad
        o += numb_param_per_model(name)
    return ok


# test hook: tests replace this factory to check the host logic without a GPU
_BACKEND_FACTORY = _default_backend


class _History:
    """Append-only [steps, W, k] store with amortised O(1) append (the reference re-allocates the
    whole history with np.concatenate every step, mcmc.py:530-534)."""

    def __init__(self, first):
        self._buf = np.array(first, dtype=np.float64, copy=True)
        self._n = self._buf.shape[0]

    def append(self, step):
        if self._n == self._buf.shape[0]:
            grown = np.zeros((max(4, 2 * self._buf.shape[0]),) + self._buf.shape[1:], dtype=np.float64)
            grown[:self._n] = self._buf[:self._n]
            self._buf = grown
        self._buf[self._n] = step[0]
        self._n += 1

    @property
    def array(self):
        return self._buf[:self._n]


class MCMC:
    """Ensemble sampler state machine (reference mcmc.py:34-567)."""

    def __init__(self, t, rv, rv_err, hparam0, kernel_name, model_par0, model_name, prior_list, numb_chains=100, flags=None, Mstar=0):
        self.t = t
        self.rv = rv
        self.rv_err = rv_err
        self.kernel_name = kernel_name
        self.model_name = model_name
        self.Mstar = Mstar
        self.flags = flags if flags is not None else None

        if isinstance(self.model_name, list):
            self.numb_models = len(self.model_name)
        elif isinstance(self.model_name, str):
            self.numb_models = 1
            self.model_name = [model_name]
        else:
            raise TypeError("model_name must be a string or a list of strings")
        self.prior_list = prior_list
        self.numb_chains = int(numb_chains)

        self.hlen = len(hparam0)
        self.plen = len(model_par0)
        W = self.numb_chains

        # number of mass columns = number of Keplerians (reference mcmc.py:134-147)
        try:
            model_par0['P'].value
            n_planets = 1
        except KeyError:
            n_planets = sum(1 for i in range(len(self.model_name)) if ('P_' + str(i)) in model_par0)
        self._n_planets = n_planets

        self.single_hp0 = [hparam0[k].value for k in hparam0.keys()]
        self.hp_err = [hparam0[k].error for k in hparam0.keys()]
        self.hp_vary = [hparam0[k].vary for k in hparam0.keys()]
        self.hparam_names = list(hparam0.keys())
        self.single_modpar0 = [model_par0[k].value for k in model_par0.keys()]
        self.modpar_err = [model_par0[k].error for k in model_par0.keys()]
        self.modpar_vary = [model_par0[k].vary for k in model_par0.keys()]
        self.modpar_names = list(model_par0.keys())

        for pr in prior_list:
            assert pr[0] in self.modpar_names or pr[0] in hparam0.keys(), "mispelt or incorrect parameter name: " + str(pr[0])

        which_model, name = [], []
        for b, mod in enumerate(self.model_name):
            for _ in range(numb_param_per_model(mod)):
                which_model.append(b)
                name.append(mod + str(b))
        self.modpar_info = [self.modpar_names, which_model, name]

        self.k_numb_param = len(self.single_hp0)
        self.mod_numb_param = len(self.single_modpar0)
        self.numb_param = self.k_numb_param + self.mod_numb_param

        # Keplerian ecc, omega -> Sk, Ck (Ford 2006), with errors (reference mcmc.py:196-198)
        for g in range(len(self.single_modpar0)):
            if self.modpar_info[0][g].startswith('ecc') and self.modpar_info[2][g].startswith(('kep', 'Kep')):
                (self.single_modpar0[g], self.single_modpar0[g + 1], self.modpar_err[g],
                 self.modpar_err[g + 1]) = aux.to_SkCk(self.single_modpar0[g], self.single_modpar0[g + 1],
                                                       self.modpar_err[g], self.modpar_err[g + 1])

        if self.numb_chains < (len(self.single_hp0) + len(self.single_modpar0)) * 2:
            # the reference *returns* a RuntimeWarning from __init__ here (mcmc.py:201-202), which
            # Python turns into this TypeError
            raise TypeError("__init__() should return None, not 'RuntimeWarning'")

        # starting positions: two np.random-consuming calls, hyper-parameters first (mcmc.py:205-206)
        self.hp0 = mcmcx.initial_pos_creator(self.single_hp0, self.hp_err, self.numb_chains)
        self.modpar0 = mcmcx.initial_pos_creator(self.single_modpar0, self.modpar_err, self.numb_chains)

        self._backend = _BACKEND_FACTORY(self.t, self.rv, self.rv_err, self.kernel_name, self.model_name,
                                         self.prior_list, self.hparam_names, self.modpar_names, self.flags)

        # initial likelihoods and masses of all walkers: one batched call (reference loop mcmc.py:221-261)
        self.logL0 = np.zeros(shape=(1, W, 1))
        self.logL0[0, :, 0] = self._evaluate(self.hp0[0], self.modpar0[0])
        self.mass0_list = self._masses(self.modpar0[0])

        self._hist_hp = _History(self.hp0)
        self._hist_mod = _History(self.modpar0)
        self._hist_logL = _History(self.logL0)
        self._hist_acc = _History(np.ones(shape=(1, W, 1)))
        self._hist_mass = _History(self.mass0_list)

        print("Initial hyper-parameter guesses: ")
        print(self.single_hp0)
        print()
        print("Initial model parameter guesses (ecc and omega are replaced by Sk and Ck): ")
        print(self.single_modpar0)
        print()
        print("Initial Log Likelihood: ", self.logL_list[0, 0, 0])
        print()
        print("Number of chains: ", self.numb_chains)
        print()

    # -- history views (same attribute names / shapes as the reference) ---------------------------
    @property
    def hparameter_list(self):
        return self._hist_hp.array

    @property
    def model_parameter_list(self):
        return self._hist_mod.array

    @property
    def logL_list(self):
        return self._hist_logL.array

    @property
    def accepted(self):
        return self._hist_acc.array

    @property
    def mass_list(self):
        return self._hist_mass.array

    # -- batched evaluation ------------------------------------------------------------------------
    def _evaluate(self, hp, modpar):
        """log L + log priors of all walkers: the reference's per-chain loop body
        (mcmc.py:421-458) as one GPU call.  A non-zero status raises what the reference's
        cho_factor would have raised, ending the run as the reference does."""
        logl, status = self._backend.logl(np.ascontiguousarray(hp), np.ascontiguousarray(modpar))
        engine.raise_for_status(status)
        return logl

    def _masses(self, modpar):
        """[1, W, n_planets] minimum masses with the reference's argument-order quirk: mass_calc
        receives [P, K, omega, ecc] but unpacks the third entry as ecc, and runs BEFORE the
        Sk,Ck -> ecc,omega conversion, so the 'eccentricity' in the formula is the chain's Ck
        (mcmc.py:438-444, auxiliary.py:242-245)."""
        W = self.numb_chains
        out = np.zeros(shape=(1, W, self._n_planets))
        if self._n_planets == 0:
            return out
        names = self.modpar_names
        cols = []
        for i in range(self._n_planets):
            sfx = "" if "P" in names else "_" + str(i)
            cols.append((names.index("P" + sfx), names.index("K" + sfx), names.index("omega" + sfx), names.index("ecc" + sfx)))
        if isinstance(self.Mstar, (int, float)) and self.Mstar == 0 and np.all(np.isfinite(modpar)) and all(
                np.all(modpar[:, iP] >= 0) and np.all(np.abs(modpar[:, iO]) <= 1) for iP, _, iO, _ in cols):
            # Mstar = 0 (the default): every mass is exactly 0.0 -- provided the formula is finite: the reference
            # yields NaN, not 0, when its "eccentricity" slot exceeds 1 or P < 0 (sqrt / cube root of a negative times 0)
            return out
        for i in range(self._n_planets):
            iP, iK, iO, iE = cols[i]
            # element by element on NumPy scalars, as the reference does: the array (SIMD) pow differs
            # from the scalar libm pow in the last bit
            for w in range(W):
                out[0, w, i] = aux.mass_calc([modpar[w, iP], modpar[w, iK], modpar[w, iO], modpar[w, iE]], self.Mstar)
        return out

    # -- proposal ----------------------------------------------------------------------------------
    def split_step(self, n_splits=2, a=2., Rstar=None, Mstar=None):
        """Stretch-move proposals for ALL walkers from the snapshot hp0/modpar0 (reference
        mcmc.py:286-403).  RNG call order kept: one random.shuffle, then per walker
        of each split one random.randint followed by np.random.rand draws (redrawn while a
        hyper-parameter is negative or parameter_check fails)."""
        W = self.numb_chains
        # same RNG consumption and result as random.shuffle on the NumPy array, without the slow
        # element-wise array swaps
        base = getattr(self, "_inds_base", None)
        if base is None or base[0] != (W, n_splits):
            base = self._inds_base = ((W, n_splits), (np.arange(W) % n_splits).tolist())
        inds = list(base[1])
        _shuffle(inds)
        inds = np.array(inds)

        self.hp = np.zeros_like(self.hp0)
        self.modpar = np.zeros_like(self.modpar0)
        self.logz = np.zeros(W)

        hp0 = self.hp0[0]
        mod0 = self.modpar0[0]
        # the reference locates each walker again by matching the first hyper-parameter and first
        # model parameter, keeping the LAST match (mcmc.py:378-381); same rule, O(W) instead of O(W^2)
        first_hp = hp0[:, 0].tolist()
        distinct = len(set(first_hp)) == W        # the usual case: every walker is its own (last) match
        keys = last_match = None
        if not distinct:
            keys = list(zip(first_hp, mod0[:, 0].tolist()))
            last_match = dict(zip(keys, range(W)))
        check_args = (Rstar, Mstar) if (Rstar is not None and Mstar is not None) else ()
        a = float(a)
        native = not check_args and self.hlen <= 64 and self.plen <= 64
        if native:
            hp0 = np.ascontiguousarray(hp0)
            mod0 = np.ascontiguousarray(mod0)
            # The Python `random` stream (one randint per chain) and the np.random stream (the uniforms) are
            # independent, and every proposal is made from the snapshot, so the chains of all splits can be
            # lined up -- split by split, in chain order -- and scanned in one native call: chain i consumes
            # the next uniform and keeps consuming while its proposal is invalid, so the numbers a chain sees
            # depend on every chain before it.  mrv_stretch_split (csrc/host_sampler.h) gathers the two walkers
            # of each chain, scans over a buffer of pre-drawn uniforms sized so that it never overdraws, and
            # scatters the proposals: the sequence of numbers each chain sees -- and the position of both
            # streams afterwards -- is exactly the reference's.
            own, partner = [], []
            for split in range(n_splits):
                in_s1 = inds == split
                idx1, idx2 = np.nonzero(in_s1)[0], np.nonzero(~in_s1)[0]
                if len(idx1) == 0:
                    np.zeros((0, 1))[0]       # an empty first set is an IndexError in the reference too
                own.append(idx1)
                partner.append(idx2[_randints0(len(idx2), len(idx1))])
            own = np.concatenate(own)
            dest = own.astype(np.int64) if distinct else np.fromiter(map(last_match.__getitem__, keys), np.int64, W)[own]
            z = self._split_native(hp0, mod0, own, np.concatenate(partner), dest, a)
            # np.log over the array runs the same per-element routine as the reference's np.log(z) on one
            # Python float (checked bit for bit); with duplicated destinations the last chain wins, as there
            self.logz[dest] = np.log(z)
            return

        for split in range(n_splits):
            in_s1 = inds == split
            idx1 = np.nonzero(in_s1)[0]
            N_chains1, N_chains2 = len(idx1), W - len(idx1)
            if N_chains1 == 0:
                np.zeros((0, 1))[0]
            # Python `random` stream: one randint per chain, in chain order (independent of np.random)
            rints = _randints0(N_chains2, N_chains1)
            positions = idx1.tolist() if distinct else [last_match[keys[i]] for i in idx1.tolist()]

            def stretch(u):
                # pure Python-float arithmetic, exactly the reference's expression (np.random.rand()
                # returns a Python float there, so `**2` is libm pow(x, 2.0), not x*x)
                return (u * (a - 1) + 1) ** 2 / a

            hp_vary = np.array(self.hp_vary, dtype=bool)
            mod_vary = np.array(self.modpar_vary, dtype=bool)
            S1_hp, S2_hp = hp0[in_s1], hp0[~in_s1]
            S1_mod, S2_mod = mod0[in_s1], mod0[~in_s1]
            Xj, Xj_mod = S2_hp[rints], S2_mod[rints]
            dX, dX_mod = S1_hp - Xj, S1_mod - Xj_mod
            z = np.empty(N_chains1)
            new_hp = np.empty_like(S1_hp)
            new_mod = np.empty_like(S1_mod)
            self._scan_python(N_chains1, Xj, dX, Xj_mod, dX_mod, stretch, check_args, z, new_hp, new_mod)
            out_hp = np.where(hp_vary, new_hp, S1_hp)
            out_mod = np.where(mod_vary, new_mod, S1_mod)
            logz = np.array([np.log(v) for v in z.tolist()])
            if len(set(positions)) == N_chains1:
                self.hp[0, positions] = out_hp
                self.modpar[0, positions] = out_mod
                self.logz[positions] = logz
            else:   # duplicated keys: later chains overwrite earlier ones, as in the reference's loop
                for c, pos in enumerate(positions):
                    self.hp[0, pos], self.modpar[0, pos], self.logz[pos] = out_hp[c], out_mod[c], logz[c]

    def _split_native(self, hp0, mod0, own, partner, dest, a):
        """One split through the C ABI host helper mrv_stretch_split (no GPU involved): gathers the two
        walkers of every chain, runs the validity-redraw scan and stores the proposals in self.hp /
        self.modpar.  Returns z per chain."""
        import ctypes
        from . import _native as nat
        lib = nat.lib()
        n_chains = len(own)
        own = np.ascontiguousarray(own, dtype=np.int64)
        partner = np.ascontiguousarray(partner, dtype=np.int64)
        cache = getattr(self, "_split_consts", None)
        if cache is None:
            kep = self._kepler_offsets()
            cache = self._split_consts = (kep, np.array(self.hp_vary, dtype=np.uint8), np.array(self.modpar_vary, dtype=np.uint8))
        kep, hp_vary, mod_vary = cache
        kep_p = kep.ctypes.data_as(nat.c_int32_p) if kep.size else None
        z = np.empty(n_chains)
        done = ctypes.c_int64(0)
        p_done = ctypes.byref(done)
        hp_out, mod_out = self.hp[0], self.modpar[0]
        # Every chain that is not finished needs at least one more uniform, so drawing exactly
        # (chains left) numbers per round never overdraws: the np.random stream ends where the
        # reference leaves it without any get_state / set_state rewind.  Draws of a chain that ran
        # dry in the middle of its redraws stay at the head of `buf` and are re-scanned.
        buf = None
        while done.value < n_chains:
            fresh = np.random.rand(n_chains - done.value)
            buf = fresh if buf is None or buf.size == 0 else np.concatenate([buf, fresh])
            used = lib.mrv_stretch_split(buf.ctypes.data, buf.size, done.value, n_chains, self.hlen, self.plen,
                                         hp0.ctypes.data, mod0.ctypes.data, own.ctypes.data, partner.ctypes.data,
                                         dest.ctypes.data, hp_vary.ctypes.data, mod_vary.ctypes.data, a, kep_p,
                                         int(kep.size), p_done, z.ctypes.data, hp_out.ctypes.data, mod_out.ctypes.data)
            if used < 0:
                raise RuntimeError("mrv_stretch_split rejected its arguments")
            buf = buf[used:]
        return z

    def _kepler_offsets(self):
        offs, o = [], 0
        for name in self.model_name:
            if name.startswith(('kepl', 'Kepl')):
                offs.append(o)
            o += numb_param_per_model(name)
        return np.array(offs, dtype=np.int32)

    def _scan_native(self, n_chains, Xj, dX, Xj_mod, dX_mod, a, z, new_hp, new_mod):
        """Validity-redraw scan of one split through the C ABI host helper (no GPU involved)."""
        import ctypes
        from . import _native as nat
        lib = nat.lib()
        Xj, dX = np.ascontiguousarray(Xj), np.ascontiguousarray(dX)
        Xj_mod, dX_mod = np.ascontiguousarray(Xj_mod), np.ascontiguousarray(dX_mod)
        kep = self._kepler_offsets()
        kep_p = kep.ctypes.data_as(nat.c_int32_p) if kep.size else None
        done = ctypes.c_int64(0)
        # Every chain that is not finished needs at least one more uniform, so drawing exactly
        # (chains left) numbers per round never overdraws: the np.random stream ends where the
        # reference leaves it without any get_state / set_state rewind.  Draws of a chain that ran
        # dry in the middle of its redraws stay at the head of `buf` and are re-scanned.
        buf = None
        nh, nm, n_kep = Xj.shape[1], Xj_mod.shape[1], int(kep.size)
        p_Xj, p_dX, p_Xjm, p_dXm = Xj.ctypes.data, dX.ctypes.data, Xj_mod.ctypes.data, dX_mod.ctypes.data
        p_z, p_hp, p_mod = z.ctypes.data, new_hp.ctypes.data, new_mod.ctypes.data
        p_done = ctypes.byref(done)
        while done.value < n_chains:
            fresh = np.random.rand(n_chains - done.value)
            buf = fresh if buf is None or buf.size == 0 else np.concatenate([buf, fresh])
            used = lib.mrv_stretch_scan(buf.ctypes.data, buf.size, done.value, n_chains, nh, nm, p_Xj, p_dX, p_Xjm, p_dXm,
                                        float(a), kep_p, n_kep, p_done, p_z, p_hp, p_mod)
            buf = buf[used:]

    def _scan_python(self, n_chains, Xj, dX, Xj_mod, dX_mod, stretch, check_args, z, new_hp, new_mod):
        """Same scan in NumPy/Python (used when Rstar and Mstar are given: the reference's orbit_check
        path raises from inside parameter_check)."""
        first = 0
        while first < n_chains:
            n = min(n_chains - first, 128)      # bounded speculation: a redraw costs <= one chunk
            state = np.random.get_state()
            u = np.random.rand(n).tolist()
            zz = np.array([stretch(x) for x in u])
            cand_hp = Xj[first:first + n] + dX[first:first + n] * zz[:, None]
            cand_mod = Xj_mod[first:first + n] + dX_mod[first:first + n] * zz[:, None]
            ok = ~(np.min(cand_hp, axis=1) < 0)
            ok &= _parameter_check_rows(cand_mod, self.model_name, check_args)
            nok = n if ok.all() else int(np.argmin(ok))
            z[first:first + nok] = zz[:nok]
            new_hp[first:first + nok] = cand_hp[:nok]
            new_mod[first:first + nok] = cand_mod[:nok]
            first += nok
            if nok == n:
                continue
            np.random.set_state(state)
            np.random.rand(nok + 1)          # replay: the accepted chains' draws + this chain's first
            while True:
                zc = stretch(np.random.rand())
                Xk_new = Xj[first] + dX[first] * zc
                Xk_new_mod = Xj_mod[first] + dX_mod[first] * zc
                if not (np.min(Xk_new) < 0) and parameter_check(Xk_new_mod, self.model_name, *check_args):
                    break
            z[first], new_hp[first], new_mod[first] = zc, Xk_new, Xk_new_mod
            first += 1

    # -- likelihood ----------------------------------------------------------------------------------
    def compute(self):
        """log L (+ priors) and masses of the proposed positions (reference mcmc.py:410-461): one
        batched GPU evaluation of all walkers."""
        W = self.numb_chains
        self.logL = np.zeros(shape=(1, W, 1))
        self.logL[0, :, 0] = self._evaluate(self.hp[0], self.modpar[0])
        self.mass = self._masses(self.modpar[0])

    # -- accept / reject -------------------------------------------------------------------------------
    def compare(self):
        """Metropolis decision per walker (reference mcmc.py:467-535): diff > 1
        accept, diff < -35 reject, otherwise accept iff random.uniform(0,1) <= exp(diff) -- the
        uniform is drawn ONLY in the middle band, in walker order.  A NaN diff matches no branch
        and records zeros, as in the reference.  The z^(n-1) factor is (as there) not applied."""
        W = self.numb_chains
        diff = self.logL[0, :, 0] - self.logL0[0, :, 0]
        accept = diff > 1
        reject = diff < -35.
        idx = np.flatnonzero((diff >= -35.) & (diff <= 1.))
        if idx.size:
            rnd = random.random     # random.uniform(0, 1) is 0 + (1 - 0) * random(): the same number, bit for bit
            ok = np.array([rnd() for _ in range(idx.size)]) <= np.exp(diff[idx])
            accept[idx[ok]] = True
            reject[idx[~ok]] = True

        # accepted rows take the proposal, rejected rows the current state, undecided (NaN) rows stay zero
        acc_col = accept[:, None]
        hp_decision = np.where(acc_col, self.hp[0], self.hp0[0])[None]
        modpar_decision = np.where(acc_col, self.modpar[0], self.modpar0[0])[None]
        logL_decision = np.where(acc_col, self.logL[0], self.logL0[0])[None]
        self.mass_decision = np.where(acc_col, self.mass[0], self.mass0_list[0])[None]
        self.acceptance_chain = accept.astype(np.float64).reshape(1, W, 1)
        undecided = ~(accept | reject)
        if undecided.any():
            for arr in (hp_decision, modpar_decision, logL_decision, self.mass_decision):
                arr[0, undecided] = 0.

        self._hist_logL.append(logL_decision)
        self._hist_mass.append(self.mass_decision)
        self._hist_acc.append(self.acceptance_chain)
        self._hist_hp.append(hp_decision)
        self._hist_mod.append(modpar_decision)

    def reset(self):
        """Accepted proposals become the new current state (reference mcmc.py:540-567); returns
        (logL_list, hparameter_list, model_parameter_list, accepted, mass_list)."""
        acc = (self.acceptance_chain[0, :, 0] == True)[:, None]  # noqa: E712
        np.copyto(self.logL0[0], self.logL[0], where=acc)
        np.copyto(self.hp0[0], self.hp[0], where=acc)
        np.copyto(self.modpar0[0], self.modpar[0], where=acc)
        np.copyto(self.mass0_list[0], self.mass[0], where=acc)
        return self.logL_list, self.hparameter_list, self.model_parameter_list, self.accepted, self.mass_list

    # -- device-resident loop (opt-in, non-parity) ----------------------------------------------------------
    def run_device(self, iterations, n_splits=2, a=2., seed=0):
        """Run ``iterations`` stretch-move iterations entirely on the GPU (``mrv_mcmc_device``,
        csrc/device_sampler.cuh) starting from the current state, and append them to the histories.

        Same algorithm as split_step / compute / compare / reset, but with a counter-based Philox
        stream on the device instead of the host's ``random`` / ``np.random``: reproducible for a
        ``seed``, NOT comparable step by step with the reference (the host loop stays the parity
        path).  Nothing crosses PCIe between iterations, so small problems are no longer bound by the
        host sampler.  Single GPU; masses are recomputed on the host
        from the recorded chains with the reference's argument order (F9), vectorised."""
        import ctypes
        from . import _native as nat
        prob = getattr(self._backend, "problem", None)
        if prob is None:
            raise RuntimeError("run_device needs the CUDA backend (no CPU fallback)")
        W, it = self.numb_chains, int(iterations)
        rng = np.random.default_rng(seed)
        base = (np.arange(W) % int(n_splits)).astype(np.uint8)
        split_all = rng.permuted(np.tile(base, (max(it, 1), 1)), axis=1)[:it]
        split_all = np.ascontiguousarray(split_all, dtype=np.uint8)
        hp_vary = np.ascontiguousarray(self.hp_vary, dtype=np.uint8)
        mp_vary = np.ascontiguousarray(self.modpar_vary, dtype=np.uint8)
        hp0 = nat.as_f64(self.hp0[0])
        mp0 = nat.as_f64(self.modpar0[0])
        hh = np.empty((it + 1, W, self.hlen))
        mh = np.empty((it + 1, W, self.plen))
        lh = np.empty((it + 1, W, 1))
        ah = np.empty((it + 1, W, 1))
        fail = np.zeros(3, dtype=np.int32)
        rc = nat.lib().mrv_mcmc_device(prob._handle, nat.dptr(hp0), nat.dptr(mp0), W, it, float(a), hp_vary.ctypes.data,
                                       mp_vary.ctypes.data, split_all.ctypes.data, int(seed) & (2**64 - 1), nat.dptr(hh),
                                       nat.dptr(mh), nat.dptr(lh), nat.dptr(ah), fail.ctypes.data_as(nat.c_int32_p))
        nat.check(rc, "mrv_mcmc_device")
        if fail[0]:
            engine.raise_for_status(np.array([fail[2]], dtype=np.int32))
        # row 0 repeats the current state: append rows 1.. to the histories and move the state
        for k in range(1, it + 1):
            self._hist_hp.append(hh[k:k + 1])
            self._hist_mod.append(mh[k:k + 1])
            self._hist_logL.append(lh[k:k + 1])
            self._hist_acc.append(ah[k:k + 1])
        acc = ah[1:, :, 0] == 1.0
        if it > 0 and acc.any():
            last = it - np.argmax(acc[::-1], axis=0)              # last accepted row per walker (valid where any)
            moved = acc.any(axis=0)
            cols = np.nonzero(moved)[0]
            self.hp0[0, cols] = hh[last[cols], cols]
            self.modpar0[0, cols] = mh[last[cols], cols]
            self.logL0[0, cols] = lh[last[cols], cols]
        if self._n_planets:
            masses = np.zeros((it, W, self._n_planets))
            names = self.modpar_names
            for i in range(self._n_planets):
                sfx = "" if "P" in names else "_" + str(i)
                iP, iK, iO = names.index("P" + sfx), names.index("K" + sfx), names.index("omega" + sfx)
                iE = names.index("ecc" + sfx)
                masses[:, :, i] = aux.mass_calc([mh[1:, :, iP], mh[1:, :, iK], mh[1:, :, iO], mh[1:, :, iE]], self.Mstar)
            for k in range(it):
                self._hist_mass.append(masses[k:k + 1])
            # the state a following split_step / compute / compare continues from: masses of the last accepted positions
            self.mass0_list = self._masses(self.modpar0[0])
        else:
            for k in range(it):
                self._hist_mass.append(np.zeros((1, W, 0)))
        return self.logL_list, self.hparameter_list, self.model_parameter_list, self.accepted, self.mass_list

    def gelman_rubin_calc(self, burn_in=200):
        """Gelman-Rubin statistic of every hyper-parameter, then every model parameter, over the walker
        chains after ``burn_in`` steps, computed on the GPU (``mrv_gelman_rubin``).

        NOT a parity function: the reference's implementation (mcmc.py:571-662) reads the
        [steps, walkers, parameters] history with its axes mislabelled and drops into pdb
        (SURVEY.md F10); this returns the statistic it intends, R = ((1 - 1/L) W + B/L) / W with
        chains = walkers, and 1.0 for a parameter that never moves (mcmc.py:632-638)."""
        from . import _native as nat
        out = []
        for hist in (self.hparameter_list, self.model_parameter_list):
            arr = nat.as_f64(hist)
            R = np.empty(arr.shape[2])
            rc = nat.lib().mrv_gelman_rubin(nat.default_device(), nat.dptr(arr), arr.shape[0], arr.shape[1], arr.shape[2],
                                            int(burn_in), nat.dptr(R))
            nat.check(rc, "mrv_gelman_rubin")
            out.append(R)
        return np.concatenate(out)


def run_MCMC(iterations, t, rv, rv_err, hparam0, kernel_name, model_param0=None, model_name=["no_model"], prior_list=[], numb_chains=None, n_splits=None, a=None, Rstar=None, Mstar=None, flags=None, plot_convergence=False, saving_folder=None, gelman_rubin_limit=1.1):
    """Run the ensemble MCMC (reference mcmc.py:676-874; same arguments, prints and return tuple:
    ``(logL_list, hparameter_list, model_parameter_list, completed_iterations)`` and additionally
    ``mass`` as the last element when ``Mstar`` is given).

    ``saving_folder`` and ``gelman_rubin_limit`` are accepted for signature compatibility and have no effect, exactly as in
    the reference whenever ``plot_convergence`` is False (they are only read inside its convergence-plot branch,
    mcmc.py:816-860, which drops into pdb -- SURVEY.md F10); ``completed_iterations`` is therefore always
    ``iterations``.  ``plot_convergence=True`` raises ``NotImplementedError`` (needs matplotlib; out of scope).  A correct
    Gelman-Rubin statistic is available separately as ``MCMC.gelman_rubin_calc`` (non-parity)."""
    if model_param0 == None:  # noqa: E711
        if model_name[0].startswith("no") or model_name[0].startswith("No"):
            model_param0 = modl.mod_create(["no_model"])
            model_param0["no"] = par.parameter(value=0., error=0., vary=False)
        else:
            raise KeyError("model parameters and model list must be provided when using a model")

    if numb_chains is not None:
        assert numb_chains > (len(model_param0) + len(hparam0)) * 2, "number of chains should be larger than the number of free parameters"

    kwargs = {}
    if flags is not None:
        kwargs["flags"] = flags
    if Mstar is not None:
        kwargs["Mstar"] = Mstar
    if numb_chains is None:
        _ = MCMC(t, rv, rv_err, hparam0, kernel_name, model_param0, model_name, prior_list, **kwargs)
        numb_chains = 100
    else:
        _ = MCMC(t, rv, rv_err, hparam0, kernel_name, model_param0, model_name, prior_list, numb_chains, **kwargs)

    print("Start Iterations")
    print()
    start = time.time()

    if plot_convergence:
        # needs gelman_rubin_calc + matplotlib, both outside the accelerated path
        raise NotImplementedError("plot_convergence=True is out of scope (reference path is broken, SURVEY.md F10)")

    step_kwargs = {}
    if n_splits is not None:
        step_kwargs["n_splits"] = n_splits
    if a is not None:
        step_kwargs["a"] = a

    for iteration in range(iterations):
        _.split_step(**step_kwargs)
        _.compute()
        if Rstar is not None and Mstar is not None:
            _.compare(Rstar=Rstar, Mstar=Mstar)      # TypeError, exactly as in the reference (F10)
        else:
            _.compare()
        logL_list, hparameter_list, model_parameter_list, accepted, mass = _.reset()
        if (iteration % 2 == 0) or iteration == iterations - 1:
            aux.printProgressBar(iteration, iterations - 1, length=50)
    completed_iterations = iterations

    print()
    print()
    print("{} iterations have been completed with {} contemporaneous chains".format(iterations, numb_chains))
    print()

    passed = int(np.count_nonzero(accepted))
    ratio = passed / (iterations * numb_chains + numb_chains)
    print("Acceptance Rate = ", ratio)
    print(" ---- %s minutes ----" % ((time.time() - start) / 60))

    if Mstar is None:
        return logL_list, hparameter_list, model_parameter_list, completed_iterations
    return logL_list, hparameter_list, model_parameter_list, completed_iterations, mass


def run_MCMC_device(iterations, t, rv, rv_err, hparam0, kernel_name, model_param0=None, model_name=["no_model"], prior_list=[], numb_chains=None, n_splits=None, a=None, Mstar=None, flags=None, seed=0):
    """``run_MCMC`` with the whole sampling loop on the GPU (``MCMC.run_device``): same arguments and
    return tuple, plus ``seed`` for the device random stream.  Opt-in and NOT step-for-step comparable
    with the reference (different random stream, same algorithm); starting positions are drawn on
    the host exactly as ``run_MCMC`` draws them."""
    if model_param0 == None:  # noqa: E711
        if model_name[0].startswith("no") or model_name[0].startswith("No"):
            model_param0 = modl.mod_create(["no_model"])
            model_param0["no"] = par.parameter(value=0., error=0., vary=False)
        else:
            raise KeyError("model parameters and model list must be provided when using a model")
    if numb_chains is not None:
        assert numb_chains > (len(model_param0) + len(hparam0)) * 2, "number of chains should be larger than the number of free parameters"
    kwargs = {}
    if flags is not None:
        kwargs["flags"] = flags
    if Mstar is not None:
        kwargs["Mstar"] = Mstar
    sampler = MCMC(t, rv, rv_err, hparam0, kernel_name, model_param0, model_name, prior_list,
                   100 if numb_chains is None else numb_chains, **kwargs)
    start = time.time()
    logL_list, hparameter_list, model_parameter_list, accepted, mass = sampler.run_device(
        iterations, n_splits=2 if n_splits is None else n_splits, a=2. if a is None else a, seed=seed)
    print("{} iterations have been completed with {} contemporaneous chains".format(iterations, sampler.numb_chains))
    print("Acceptance Rate = ", int(np.count_nonzero(accepted)) / (iterations * sampler.numb_chains + sampler.numb_chains))
    print(" ---- %s minutes ----" % ((time.time() - start) / 60))
    if Mstar is None:
        return logL_list, hparameter_list, model_parameter_list, iterations
    return logL_list, hparameter_list, model_parameter_list, iterations, mass
