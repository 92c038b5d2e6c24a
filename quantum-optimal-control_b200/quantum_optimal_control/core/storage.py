"""Run files with the reference's dataset schema.

The reference writes HDF5 through ``H5File.add`` / ``H5File.append`` (helper_functions/
data_management.py:109-156): ``add`` stores a value once, ``append`` grows a dataset along a new
leading axis, one slice per save.  Datasets (main_grape/grape.py:55-87, core/run_session.py:129-138,
core/analysis.py:31-33,62-65,95-99, core/system_parameters.py:189-191,233-236):

    inputs        H0 Hops Hnames U total_time steps states_concerned_list use_gpu sparse_* maxA
                  initial_guess method convergence/* reg_coeffs/* dressed_info/*
    setup         initial_vectors_c taylor_terms taylor_scaling
    per save      error reg_error uks iteration run_time unitary_scale
    per evol save final_state (real-embedded 2n x 2n)  inter_vecs_raw_real/_imag
                  inter_vecs_mag_squared inter_vecs_real inter_vecs_imag   ([m, n, T+1] each)
    end           wall_clock_time

h5py is optional in this image: with it the file is a real ``.h5``; without it the same keys go
into a NumPy ``.npz`` (``group/key`` for the three input dictionaries).
"""
import os

import numpy as np

try:                                        # pragma: no cover - depends on the image
    import h5py
except Exception:                           # noqa: BLE001
    h5py = None


def new_run_file(data_path, file_name):
    """NNNNN_<name>.h5 with the first free 5-digit prefix (main_grape/grape.py:44-50).  The name is RESERVED
    atomically (the empty file is created with O_CREAT | O_EXCL) so that concurrent processes -- one per GPU in
    ``core.population`` -- never resolve the same path."""
    ext = ".h5" if h5py is not None else ".npz"
    num = 0
    while True:
        path = os.path.join(data_path, str(num).zfill(5) + "_" + file_name + ext)
        try:
            os.close(os.open(path, os.O_CREAT | os.O_EXCL | os.O_WRONLY))
            return path
        except FileExistsError:
            num += 1


class RunFile:
    """``add`` / ``append`` with H5File semantics on top of h5py or an in-memory dict flushed to .npz."""

    def __init__(self, path):
        self.path = path
        self._once, self._series = {}, {}
        if os.path.exists(path) and os.path.getsize(path) == 0:        # name reserved by new_run_file, nothing written yet
            if h5py is not None:
                os.remove(path)
        elif h5py is None and os.path.exists(path):
            with np.load(path, allow_pickle=True) as f:
                self._once = {k: f[k] for k in f.files}

    def add(self, key, data):
        if data is None:
            return
        if h5py is not None:
            with h5py.File(self.path, 'a') as hf:
                if key in hf:
                    del hf[key]
                try:
                    hf.create_dataset(key, data=data)
                except TypeError:
                    hf.create_dataset(key, data=np.array(data, dtype='S'))
            return
        self._once[key] = np.asarray(data)
        self._flush()

    def add_group(self, group, mapping):
        for k, v in (mapping or {}).items():
            self.add(group + '/' + k, v)

    def append(self, key, data):
        data = np.asarray(data)
        if h5py is not None:
            with h5py.File(self.path, 'a') as hf:
                if key not in hf:
                    hf.create_dataset(key, shape=(1,) + data.shape, maxshape=(None,) * (data.ndim + 1), dtype=data.dtype)
                    hf[key][0] = data
                else:
                    ds = hf[key]
                    ds.resize((ds.shape[0] + 1,) + ds.shape[1:])
                    ds[-1] = data
            return
        self._series.setdefault(key, []).append(data)
        self._flush()

    def _flush(self):
        out = dict(self._once)
        out.update({k: np.stack(v) for k, v in self._series.items()})
        tmp = self.path + ".tmp.npz"
        np.savez(tmp, **out)
        os.replace(tmp, self.path)


def save_inputs(run, inputs, convergence, reg_coeffs, dressed_info):
    """main_grape/grape.py:55-87."""
    for k, v in inputs.items():
        run.add(k, v)
    run.add_group('convergence', convergence)
    run.add_group('reg_coeffs', reg_coeffs)
    run.add_group('dressed_info', dressed_info)
