"""Result containers of the power-spectrum path -- host-side mirror of the reference `vpower/spctrm.py`.

Same names, arguments and return conventions as the reference (file:line cited per item) so that spectra
produced by the B200 path can be added to / compared with spectra pickled by the reference.  No GPU work
happens here ("boundary only").  Plotting helpers (`peek`, `plot`,
`high_pass_filter_2d`) are out of scope.
"""
import os
import pickle

import numpy as np


class PowerSpectrum:
    """(k, P, Psum, Nsample) table.  Reference: vpower/spctrm.py:55-245."""

    def __init__(self, Pk, m=0, beta=np.array([-1, -1, -1])) -> None:
        self.k = Pk[:, 0]
        self.P = Pk[:, 1]
        self.Psum = Pk[:, 2]
        self.Nsample = Pk[:, 3]
        self.m = m            # folding factor (0: not folded)
        self.beta = beta      # residue class of a folded sub-spectrum
        self.check_alignment()

    def data(self):
        """[n,4] stack (k, P, Psum, Nsample).  spctrm.py:68-71."""
        return np.stack([self.k, self.P, self.Psum, self.Nsample], axis=1)

    def subtract_shot_noise(self, Lbox, Np) -> None:
        """spctrm.py:73-76."""
        self.P -= Lbox ** 3 / Np
        self.P[self.P < 0] = 0

    def __len__(self):
        """Number of points; raises if the four columns are not aligned.  spctrm.py:78-91."""
        n = len(self.k)
        for name in ("P", "Psum", "Nsample"):
            if n != len(getattr(self, name)):
                raise Exception("k and {} have different length.".format(name))
        return n

    check_alignment = __len__

    def kmin(self):
        return np.min(self.k)

    def kmax(self):
        return np.max(self.k)

    def kres(self):
        """spctrm.py:99-102."""
        return (self.kmax() - self.kmin()) / (self.__len__() - 1)

    def Lbox(self):
        return 2 * np.pi / self.kmin()

    def energy(self):
        """Direct integral sum(P[:-1] * dk).  spctrm.py:108-113."""
        dk = self.k[1:] - self.k[:-1]
        return np.sum(self.P[:-1] * dk)

    def copy(self):
        return PowerSpectrum(self.data(), self.m, self.beta)

    def _same_length(self, other):
        if self.__len__() != len(other):
            raise Exception("Spectra has different length therefore cannot be combined directly.")

    def add(self, spctrm) -> None:
        """Accumulate shell sums and counts, re-derive P = Psum/Nsample * 4 pi k^2.  spctrm.py:118-126."""
        self._same_length(spctrm)
        self.Psum += spctrm.Psum
        self.Nsample += spctrm.Nsample
        self.P = self.Psum / self.Nsample * (4 * np.pi * self.k ** 2)

    def remove(self, spctrm) -> None:
        """spctrm.py:128-140."""
        self._same_length(spctrm)
        self.Psum -= spctrm.Psum
        self.Nsample -= spctrm.Nsample
        if (self.Nsample < 0).any():
            raise ValueError("Nsample is less than zero.")
        if (self.Psum < 0).any():
            raise ValueError("Psum is less than zero.")
        self.P = self.Psum / self.Nsample * (4 * np.pi * self.k ** 2)

    def append(self, spctrm) -> None:
        """Splice a coarser spectrum onto the high-k end.  spctrm.py:142-165 (the reference rebinds `self`
        at the end, so -- as there -- the receiver is left unchanged; the merged table is returned here)."""
        w = spctrm.kres()
        keep = self.k < spctrm.k[0]
        full = PowerSpectrum(np.concatenate((self.data()[keep], spctrm.data())))
        for k in spctrm.k[spctrm.k < self.k[-1]]:
            sel = ((k - w / 2) <= self.k) & (self.k < (k + w / 2))
            full.Psum[np.where(full.k == k)] += np.sum(self.Psum[sel])
            full.Nsample[np.where(full.k == k)] += np.sum(self.Nsample[sel])
        ok = np.where(full.Psum > 0)
        full.P[ok] = full.Psum[ok] / full.Nsample[ok] * (4 * np.pi * full.k[ok] ** 2)
        full.check_alignment()
        return full

    def index(self):
        """Log-log slope of the non-zero part.  spctrm.py:167-174."""
        sel = self.P > 0
        power, _ = np.polyfit(np.log10(self.k[sel]), np.log10(self.P[sel]), 1)
        return power

    def _filename(self, run_output_dir, beta):
        if beta is None or (np.asarray(beta) == np.array([-1, -1, -1])).all():
            return os.path.join(run_output_dir, "full_spctrm.pkl")
        return os.path.join(run_output_dir, "sub_spctrm_b{}{}{}.pkl".format(*beta))

    def save(self, run_output_dir) -> None:
        """Pickle under run_output_dir with the reference's file names.  spctrm.py:224-233."""
        with open(self._filename(run_output_dir, self.beta), "wb") as f:
            pickle.dump(self, f)

    @staticmethod
    def load(run_output_dir, beta=None):
        """spctrm.py:235-245."""
        name = os.path.join(run_output_dir, "full_spctrm.pkl") if beta is None else \
            os.path.join(run_output_dir, "sub_spctrm_b{}{}{}.pkl".format(*beta))
        with open(name, "rb") as f:
            return pickle.load(f)


class SpectrumList:
    """Sub-spectra of a folded transform, addressed by beta.  Reference: spctrm.py:252-315."""

    def __init__(self, spctrm_list):
        self.list = spctrm_list
        self.m = spctrm_list[0].m

    def __len__(self):
        return len(self.list)

    def __getitem__(self, beta) -> PowerSpectrum:
        for s in self.list:
            if (s.beta == beta).all():
                return s
        raise Exception("No spectrum in the list with beta = {}".format(beta))

    def __setitem__(self, beta, spctrm) -> None:
        for i, s in enumerate(self.list):
            if (s.beta == beta).all():
                self.list[i] = spctrm
                return
        self.list.append(spctrm)

    def __iter__(self):
        return iter(self.list)

    def combine_all(self) -> PowerSpectrum:
        """spctrm.py:277-282."""
        combined = empty_spectrum_like(self.list[0])
        for s in self.list:
            combined.add(s)
        return combined

    def combine_from_beta_sequence(self, beta_sequence=None) -> PowerSpectrum:
        """spctrm.py:284-291."""
        if beta_sequence is None:
            beta_sequence = init_beta_space(m=self.m)
        combined = empty_spectrum_like(self.list[0])
        for beta in beta_sequence:
            combined.add(self[beta])
        return combined

    def append(self, spctrm) -> None:
        self.list.append(spctrm)

    def save(self, run_output_dir) -> None:
        with open(os.path.join(run_output_dir, "spctrm_list.pkl"), "wb") as f:
            pickle.dump(self, f)

    @staticmethod
    def load(run_output_dir):
        """Collect every sub_spctrm_b???.pkl of a directory.  spctrm.py:302-315."""
        found = None
        for name in sorted(os.listdir(run_output_dir)):
            if name.startswith("sub_spctrm_b"):
                beta = np.array([int(x) for x in name[-7:-4]])
                s = PowerSpectrum.load(run_output_dir, beta=beta)
                if found is None:
                    found = SpectrumList([s])
                else:
                    found.append(s)
        return found


def relative_diff(spctrm1, spctrm2, mode="max") -> float:
    """max / rms / sum relative difference of two spectra.  spctrm.py:321-346."""
    if len(spctrm1) != len(spctrm2):
        raise Exception("Spectra has different length therefore cannot be compared directly.")
    P1 = spctrm1.P
    P1[np.isnan(P1)] = 0
    P1[P1 == 0] = 1e-10
    P2 = spctrm2.P
    P2[np.isnan(P2)] = 0
    if mode == "mean":
        return np.mean(((P1 - P2) / P1) ** 2) ** 0.5
    if mode == "max":
        return np.max(abs(P1 - P2) / P1)
    if mode == "sum":
        return np.sum(((P1 - P2) / P1) ** 2) ** 0.5
    raise Exception("Mode not recognized. Use 'mean' or 'max'.")


def empty_spectrum_like(spctrm, keep_m=False, keep_beta=False) -> PowerSpectrum:
    """spctrm.py:349-356."""
    k = spctrm.k
    z = np.zeros_like(k)
    return PowerSpectrum(np.column_stack((k, z, z, z)), m=spctrm.m if keep_m else 0,
                         beta=spctrm.beta if keep_beta else np.array([-1, -1, -1]))


def load_spectrum(filename) -> PowerSpectrum:
    """spctrm.py:359-363."""
    with open(filename, "rb") as f:
        return pickle.load(f)


def init_beta_space(m):
    """All residues (m^3, 3).  spctrm.py:366-372."""
    b = np.arange(0, m)
    return np.array(np.meshgrid(b, b, b, indexing="ij")).T.reshape(-1, 3)


def random_beta_sequence(m, seed=1):
    """spctrm.py:375-380 (the reference discards the permutation it draws; reproduced)."""
    np.random.seed(seed)
    beta_space = init_beta_space(m)
    np.random.permutation(beta_space)
    return beta_space
