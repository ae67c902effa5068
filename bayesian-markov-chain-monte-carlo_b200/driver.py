"""``RSF`` -- experiment-driver facade with the reference's surface (RSF.py:59-1046).

Same constructor, attributes (``num_dc``, ``dc_list``, ``qstart``, ``qpriors`` ...; ``model``,
``data`` and ``format`` are set by the caller as in main.py:162-163,300) and methods.  What changes:

  * ``generate_time_series`` solves the whole ``dc_list`` sweep in ONE batched launch
    (RSF.py:358-369 loops serially); the noise draws consume the global NumPy generator in the
    reference's order, so ``np.random.seed`` reproduces the reference's data layout.
  * ``prepare_data`` uses the JSON codec explicitly.  The reference imports the MySQL functions over
    the JSON ones (RSF.py:3-4), so its own ``prepare_data('json')`` raises TypeError (quirk q1); the
    intended JSON round trip is what is implemented.  ``format = 'mysql'`` needs a database server and
    is out of scope.
  * ``inference`` is wrapped by ``measure_execution_time`` and therefore returns elapsed seconds, not
    the samples (quirk q11); results are kept in ``self.results[dc]``.
  * plotting / animation (matplotlib, ffmpeg) is out of scope: ``plot_*`` are no-ops unless
    ``plotfigs`` is set and matplotlib is importable; the KDE they draw is available from
    ``posterior.gaussian_kde_pdf``.
"""
import time

import numpy as np

from .ndarray_json import load_object, save_object
from .sampler import MCMC


def measure_execution_time(func):
    """RSF.py:7-55: returns the elapsed wall time and DISCARDS the wrapped function's result."""
    def wrapper(*args, **kwargs):
        start_time = time.time()
        func(*args, **kwargs)
        return time.time() - start_time
    return wrapper


class RSF:
    def __init__(self, number_slip_values=1, lowest_slip_value=1.0, largest_slip_value=1000.0, qstart=10.0,
                 qpriors=["Uniform", 0.0, 10000.0], reduction=False, plotfigs=False):
        # RSF.py:250-256
        self.num_dc = number_slip_values
        self.dc_list = np.linspace(lowest_slip_value, largest_slip_value, self.num_dc)
        self.num_features = 2
        self.plotfigs = plotfigs
        self.qstart = qstart
        self.qpriors = qpriors
        self.reduction = reduction
        # extensions
        self.mcmc_kwargs = {}          # forwarded to MCMC (n_chains, seed, ...)
        self.results = {}              # dc -> dict(samples, std2, acceptance_ratio)

    # -- data ------------------------------------------------------------
    def generate_time_series(self):
        """RSF.py:260-371: noisy acceleration series for every Dc, concatenated."""
        n = self.model.num_tsteps
        out = self.model.evaluate_batch(self.dc_list, want_acc=True, want_t=self.plotfigs)
        acc_all = out["acc"].t().contiguous().cpu().numpy()                 # [num_dc, n_out]
        if acc_all.shape[1] != n:
            # the reference fails here too (quirk q8: floor((T1-T0)/dt) came out as N-1)
            raise ValueError(f"could not broadcast input array from shape ({acc_all.shape[1]},) into shape ({n},)")
        acc_appended_noise = np.zeros(len(self.dc_list) * n)
        for index, dc_value in enumerate(self.dc_list):
            self.model.Dc = dc_value
            acc = acc_all[index]
            acc_noise = acc + 1.0 * np.abs(acc) * np.random.randn(acc.shape[0])     # RateStateModel.py:392
            if self.plotfigs:
                self.plot_time_series(out["t"][:, index].cpu().numpy(), acc)
            acc_appended_noise[index * n:(index + 1) * n] = acc_noise
        return acc_appended_noise

    def prepare_data(self, data):
        """RSF.py:477-604 (JSON round trip through ``data.json``)."""
        if self.format == 'json':
            self.lstm_file = 'model_lstm.json'
            self.data_file = 'data.json'
            save_object(data, self.data_file)
            data = load_object(self.data_file)
        elif self.format == 'mysql':
            raise NotImplementedError("format='mysql' needs a MySQL server (mysql_save_load.py); out of scope")
        return data

    # -- plots (out of scope; kept as guarded no-ops) --------------------------
    def plot_time_series(self, time_values, acceleration):
        if not self.plotfigs:
            return
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            return
        plt.figure()
        plt.title(rf'$d_c$={self.model.Dc} $\mu m$ RSF solution')
        plt.plot(time_values, acceleration, linewidth=1.0, label='True')
        plt.xlabel('Time (sec)')
        plt.ylabel(r'Acceleration $(\mu m/s^2)$')

    def plot_dist(self, qparams, dc):
        if not self.plotfigs:
            return
        from .posterior import gaussian_kde_pdf
        self.results.setdefault(dc, {})["kde"] = gaussian_kde_pdf(qparams[0, :])

    # -- inference -----------------------------------------------------------
    def perform_sampling_and_plotting(self, data, dc, nsamples, model_lstm):
        """RSF.py:748-900: slice the series of this Dc and run the sampler on it."""
        index = np.where(self.dc_list == dc)[0][0] if dc in self.dc_list else -1
        if index == -1:
            print(f"Error: dc value {dc} not found in dc_list.")
            return
        start = index * self.model.num_tsteps
        end = start + self.model.num_tsteps
        noisy_data = data[start:end]
        print(f'--- Dc is {dc} ---')
        mcmc_obj = MCMC(self.model, noisy_data, dc, self.qpriors, self.qstart, lstm_model=model_lstm,
                        nsamples=nsamples, **self.mcmc_kwargs)
        qparams = mcmc_obj.sample(False)            # the reference passes True (mp4 animation, out of scope)
        self.results[dc] = {"samples": qparams, "std2": mcmc_obj.std2,
                            "acceptance_ratio": mcmc_obj.acceptance_ratio}
        self.plot_dist(qparams if qparams.ndim == 2 else qparams[0], dc)

    @measure_execution_time
    def inference(self, nsamples):
        """RSF.py:902-1046; returns elapsed seconds through the decorator."""
        data = self.prepare_data(self.data)
        for dc in self.dc_list:
            self.perform_sampling_and_plotting(data, dc, nsamples, None)
        return
