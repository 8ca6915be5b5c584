"""Host mirror of the reference's ``magpy_rv.saving`` (src/magpy_rv/saving.py:29-431): dump the inputs,
initial conditions, priors, posterior chains and final parameter values of a run to a folder, with the
same file names and contents.

Nothing here is on the hot path except the two single ``GPLikelihood.LogL`` evaluations (initial and
final log-likelihood), which run on the GPU through ``gp_likelihood``.  The reference's control flow is
a sequence of ``try / except: continue`` blocks whose observable effect is WHICH lines reach the files
when optional arguments are missing; the helpers below keep that effect (each comment names the
reference lines it reproduces) without reproducing the code.  ``offset_subtract`` lives in the
reference's ``plotting`` module (plotting.py:34-62), which cannot be imported without matplotlib; it is
a pure array function and is provided here.
"""
import os
import pickle as pk

import numpy as np

from . import auxiliary as aux
from . import gp_likelihood as gp
from . import models as mod
from . import parameters as par
from .mcmc_aux import get_model


def offset_subtract(rv, flags, offsets):
    """RVs with the per-instrument offsets removed (reference plotting.py:34-62): flag 0 subtracts
    nothing, flag a + 1 subtracts ``offsets[a]``; a flag with no matching offset contributes no entry,
    so the final subtraction fails on the length mismatch exactly as the reference's does."""
    table = {float(a + 1): offsets[a] for a in range(len(offsets))}
    y_off = []
    for f in flags:
        if f == 0.:
            y_off.append(0.)
        elif float(f) in table:
            y_off.append(table[float(f)])
    return rv - y_off


def _offsets_of(init_param, model_list):
    """Initial offset values in model order: 'offset' for a single Offset model, else every
    'offset_i' that exists (saving.py:90-101)."""
    try:
        return [init_param['offset'].value]
    except Exception:
        found = []
        for i in range(len(model_list)):
            try:
                found.append(init_param['offset_' + str(i)].value)
            except Exception:
                continue
        return found


def _mass_names_from_params(init_param, model_list):
    """'mass' for a single Keplerian, 'mass_i' for every model i that has a 'P_i' (saving.py:210-219)."""
    try:
        init_param['P'].value
        return ['mass']
    except Exception:
        names = []
        for i in range(len(model_list)):
            try:
                init_param['P_' + str(i)].value
                names.append('mass_' + str(i))
            except Exception:
                continue
        return names


def _dump_chain(folder, pkl_name, names, chain, burnin, txt_pattern="{}_posteriors.txt"):
    """One pickle of ``chain[burnin:]`` plus one text table [iterations, chains] per named column
    (saving.py:152-200)."""
    with open(os.path.join(folder, pkl_name), "wb") as fh:
        pk.dump(chain[burnin:], fh)
    for col, name in enumerate(names):
        np.savetxt(os.path.join(folder, txt_pattern.format(name)), chain[burnin:, :, col])


def _write_value_with_errors(fh, name, value, upper, lower, ups, downs):
    """One entry of final_parameter_values.txt: the value, then '+<upper>' and '<lower>' when they can be
    formed.  ``upper`` / ``lower`` are callables so that a failure (e.g. no error list given) happens at the
    same point of the line as in the reference: the '+' has already been written when the plain
    subtraction fails (saving.py:240-262, 283-322)."""
    fh.write("\n{}:\n".format(name))
    fh.write(value.__str__())
    try:
        up = upper(fh)
        ups.append(up)
        fh.write(up.__str__())
    except Exception:
        return
    try:
        down = lower(fh)
        downs.append(down)
        fh.write(down.__str__())
    except Exception:
        return


def save(folder_name, rv, time, rv_err, model_list=None, init_hparam=None, kernel=None, init_param=None, prior_list=[], fin_hparam_post=None, fin_param_post=None, logl_chain=None, masses=None, fin_param_values=None, fin_param_erru=None, fin_param_errd=None, flags=None, Mstar=None, fin_to_skck=False, burnin=0):
    """Same arguments and output files as the reference's ``saving.save`` (saving.py:29)."""
    if not os.path.exists(folder_name):
        os.mkdir(folder_name)
    path = lambda name: os.path.join(folder_name, name)   # noqa: E731

    logL = None     # like the reference's local of that name: set by the initial evaluation, reused by the table

    # data (saving.py:84-111)
    if flags is not None:
        np.savetxt(path("rv_data_offset_sub.txt"), offset_subtract(rv, flags, _offsets_of(init_param, model_list)))
    np.savetxt(path("time_data.txt"), time)
    np.savetxt(path("rv_data.txt"), rv)
    np.savetxt(path("rv_err.txt"), rv_err)

    # initial conditions (saving.py:113-141).  The reference only opens the file when a kernel is given and
    # then writes to / closes it unconditionally, i.e. it needs a kernel.
    if kernel is None:
        raise UnboundLocalError("cannot access local variable 'initial_cond_file' where it is not associated with a value")
    with open(path("initial_conditions.txt"), "w+") as fh:
        fh.write("\nKernel Name:\n" + kernel)
        fh.write("\nInitial Hyperparameters:\n" + init_hparam.__str__())
        if model_list is not None:
            fh.write("\nModel List:\n" + model_list.__str__())
            fh.write("\nInitial Parameters:\n" + init_param.__str__())
        if Mstar is not None:
            fh.write("\nHost Star Mass:\n" + Mstar.__str__() + " Solar Masses")
            # the initial log-likelihood is only recorded together with the stellar mass (saving.py:128-139)
            model_y = get_model(model_list, time, init_param, to_ecc=False, flags=flags)
            logL = gp.GPLikelihood(time, rv, rv_err, init_hparam, kernel, model_y, init_param).LogL(prior_list)
            fh.write("\nInitial LogL:\n" + logL.__str__())

    with open(path("priors.txt"), "w+") as fh:
        fh.write("\nPriors:\n" + prior_list.__str__())

    assert type(burnin) == int, "burnin should be an integer"   # noqa: E721

    # posterior chains (saving.py:150-200)
    if fin_hparam_post is not None:
        _dump_chain(folder_name, "hparam_posteriors.pkl", aux.hparam_names(kernel, plotting=False), fin_hparam_post, burnin)
    if fin_param_post is not None:
        _dump_chain(folder_name, "param_posteriors.pkl", aux.model_param_names(model_list, SkCk=True, plotting=False),
                    fin_param_post, burnin)
    if logl_chain is not None:
        np.savetxt(path("logL_posteriors.txt"), logl_chain[burnin:, :, 0])
        with open(path("logl_posteriors.pkl"), "wb") as fh:
            pk.dump(logl_chain[burnin:], fh)
    if masses is not None:
        n_mass = len(masses[0, 0, :])
        names = ["mass"] if (n_mass == 1 and len(model_list) == 1) else ["mass_{}".format(i) for i in range(n_mass)]
        _dump_chain(folder_name, "mass_posteriors.pkl", names, masses, burnin)

    if fin_param_values is None:
        return

    # ---- final values (saving.py:203-431) ----------------------------------------------------------------
    def all_names(plotting):
        """hyper-parameters + model parameters (+ masses when more values than names were given); only the
        hyper-parameter names when there is no usable model list (saving.py:204-225, 386-404)."""
        hnames = aux.hparam_names(kernel, plotting=plotting)
        try:
            pnames = aux.model_param_names(model_list, SkCk=fin_to_skck, plotting=plotting)
            extra = _mass_names_from_params(init_param, model_list) if len(fin_param_values) > len(hnames) + len(pnames) else []
            return hnames, hnames + pnames + extra
        except Exception:
            return hnames, hnames

    hparams, param_list = all_names(plotting=False)
    value_list, erru_list, errd_list = [], [], []
    with open(path("final_parameter_values.txt"), "w+") as fh:
        state = {}          # ecc / omega results carried from an 'ecc' entry to the 'omega' entry that follows

        # the reference handles exactly the two booleans (`if fin_to_skck is True:` ... `if ... is False:`)
        for N, name in enumerate(param_list if (fin_to_skck is True or fin_to_skck is False) else []):
            if fin_to_skck is True or not name.startswith(('ecc', 'omega')):
                # plain entry: value as given; '+' is written before the subtraction is attempted
                value = fin_param_values[N]

                def upper(f, N=N, value=value):
                    f.write("\n+")
                    return fin_param_erru[N] - value

                def lower(f, N=N, value=value):
                    f.write("\n")
                    return fin_param_errd[N] - value
            elif name.startswith('ecc'):
                # Sk, Ck -> ecc, omega for the value and, separately, for the two error bounds
                # (saving.py:266-272, 284-291, 305-312): the bound lists are indexed BEFORE anything is written
                state['ecc'], state['omega'] = aux.to_ecc(fin_param_values[N], fin_param_values[N + 1])
                value = state['ecc']

                def upper(f, N=N):
                    e_up, state['omega_up'] = aux.to_ecc(fin_param_erru[N], fin_param_erru[N + 1])
                    f.write("\n+")
                    return e_up - state['ecc']

                def lower(f, N=N):
                    e_dn, state['omega_dn'] = aux.to_ecc(fin_param_errd[N], fin_param_errd[N + 1])
                    f.write("\n")
                    return e_dn - state['ecc']
            else:
                value = state['omega']                      # NameError in the reference without a preceding 'ecc'

                def upper(f):
                    f.write("\n+")
                    return state['omega_up'] - state['omega']

                def lower(f):
                    f.write("\n")
                    return state['omega_dn'] - state['omega']
            value_list.append(value)
            _write_value_with_errors(fh, name, value, upper, lower, erru_list, errd_list)

    value_list = ['%.3f' % v for v in value_list]
    erru_list = ['%.3f' % v for v in erru_list]
    errd_list = ['%.3f' % v for v in errd_list]

    # final log-likelihood from the final values with their upper errors (saving.py:333-383); any failure
    # silently skips final_info.txt, as in the reference -- and leaves `logL` at the INITIAL value when one was
    # computed, which is what the table below then sees.  (For the QuasiPer kernels this block always fails
    # there and here: hparam_names spells 'gp_perlegth', so the rebuilt dictionary keeps a placeholder string.)
    iterations = num_chains = None
    have_info = False
    try:
        hnames = aux.hparam_names(kernel, plotting=False)
        pnames = aux.model_param_names(model_list, SkCk=False, plotting=False)
        new_hparam = par.par_create(kernel)
        for N, name in enumerate(hnames):
            new_hparam[name] = par.parameter(fin_param_values[N], fin_param_erru[N] - fin_param_values[N], True)
        new_param = mod.mod_create(model_list)
        nh = len(hparams)
        conv = {}
        for N, name in enumerate(pnames):
            if name.startswith('ecc'):
                conv['ecc'], conv['omega'] = aux.to_ecc(fin_param_values[N + nh], fin_param_values[N + nh + 1])
                e_up, conv['omega_up'] = aux.to_ecc(fin_param_erru[N + nh], fin_param_erru[N + nh + 1])
                new_param[name] = par.parameter(conv['ecc'], e_up - conv['ecc'], True)
            elif name.startswith('omega'):
                new_param[name] = par.parameter(conv['omega'], conv['omega_up'] - conv['omega'], True)
            else:
                v = fin_param_values[N + nh]
                new_param[name] = par.parameter(v, fin_param_erru[N + nh] - v, True)
        model_y = get_model(model_list, time, new_param, to_ecc=False, flags=flags)
        logL = gp.GPLikelihood(time, rv, rv_err, new_hparam, kernel, model_y, new_param).LogL(prior_list)
        with open(path("final_info.txt"), "w+") as fh:
            fh.write("\nFinal LogL:\n" + logL.__str__())
            try:
                num_chains = len(fin_hparam_post[0, :, 0])
                iterations = len(fin_hparam_post[:, 0, 0]) - 1
                fh.write("\nNumber of Chains:\n" + num_chains.__str__())
                fh.write("\nCompleted Iterations:\n" + iterations.__str__())
                have_info = True
            except Exception:
                pass
    except Exception:
        pass

    # LaTeX table (saving.py:385-431): with the run summary rows when the final log-likelihood and the chain
    # shape are known, else the parameter rows alone; silently nothing if even that fails
    _, table_names = all_names(plotting=True)
    import pandas as pd

    def latex(names, values, ups, downs):
        tab = np.rot90(np.array([downs, ups, values, names]), 1, axes=(1, 0))
        return pd.DataFrame(tab).to_latex(index=False, header=['Parameter', 'Value', 'Error Up', 'Error Down'])

    with open(path("final_parameter_table.txt"), "w+") as fh:
        try:
            if logL is None:
                fh.write(latex(table_names, value_list, erru_list, errd_list).__str__())
            elif have_info:
                fh.write(latex(table_names + ['Final LogL', 'Num Iterations', 'Num Chains'],
                               value_list + ["%.3f" % logL, iterations, num_chains], erru_list + [0, 0, 0],
                               errd_list + [0, 0, 0]).__str__())
            # final log L known but no chain shape: the reference has already extended the name column when
            # it fails on the missing counts, the fallback table is ragged and nothing is written (saving.py:406-431)
        except Exception:
            pass
