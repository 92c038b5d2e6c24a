"""Run files.  The reference writes HDF5 through ``H5File`` (helper_functions/data_management.py,
main_grape/grape.py:44-87, core/run_session.py:129-138); h5py is optional here: when it is missing
the same keys go into a NumPy ``.npz`` next to where the ``.h5`` would have been."""
import os

import numpy as np

try:                                        # pragma: no cover - depends on the image
    import h5py
except Exception:                           # noqa: BLE001
    h5py = None


def new_run_file(data_path, file_name):
    """NNNNN_<name>.h5 with the first free 5-digit prefix (main_grape/grape.py:44-50)."""
    ext = ".h5" if h5py is not None else ".npz"
    num = 0
    while os.path.exists(os.path.join(data_path, str(num).zfill(5) + "_" + file_name + ext)):
        num += 1
    return os.path.join(data_path, str(num).zfill(5) + "_" + file_name + ext)


def _flatten(inputs, convergence, reg_coeffs, dressed_info):
    out = {}
    for k, v in inputs.items():
        if v is not None:
            out[k] = v
    for group, d in (('convergence', convergence), ('reg_coeffs', reg_coeffs), ('dressed_info', dressed_info)):
        for k, v in (d or {}).items():
            out[group + '/' + k] = v
    return out


def _merge(path, new):
    if h5py is not None:
        with h5py.File(path, 'a') as hf:
            for k, v in new.items():
                if k in hf:
                    del hf[k]
                try:
                    hf.create_dataset(k, data=v)
                except TypeError:
                    hf.create_dataset(k, data=np.array(v, dtype='S'))
        return
    old = {}
    if os.path.exists(path):
        with np.load(path, allow_pickle=True) as f:
            old = {k: f[k] for k in f.files}
    old.update({k: np.asarray(v) for k, v in new.items()})
    tmp = path + ".tmp.npz"
    np.savez(tmp, **old)
    os.replace(tmp, path)


def save_inputs(path, inputs, convergence, reg_coeffs, dressed_info):
    _merge(path, _flatten(inputs, convergence, reg_coeffs, dressed_info))


def save_scalar(path, key, value):
    _merge(path, {key: np.array(value)})


def save_results(path, session, sys_para, wall_clock_time):
    """Final values of the datasets the reference appends every update_step
    (core/run_session.py:129-138) plus taylor_terms / wall_clock_time."""
    h = session.history
    _merge(path, {
        'error': np.array([x[0] for x in h]), 'reg_error': np.array([x[1] for x in h]),
        'unitary_scale': np.array([x[3] for x in h]), 'uks': np.asarray(session.uks),
        'iteration': np.array(session.iterations), 'taylor_terms': np.array(sys_para.exp_terms),
        'taylor_scaling': np.array(sys_para.scaling), 'wall_clock_time': np.array(wall_clock_time),
        'final_state': np.asarray(session.Uf) if not sys_para.state_transfer else np.zeros(0),
        'initial_vectors_c': np.array(sys_para.initial_vectors_c),
    })
