"""maua/audiovisual/audioreactive/util.py: inspection helpers.  ``info`` is kept; the matplotlib / librosa.display plots
(plot_signals, plot_spectra, plot_audio, plot_chroma_comparison) are debugging aids outside the render path: they are
accepted and do nothing, so patch files that call them keep running."""


def info(arr):
    """Shape and min / mean / max of (lists of) arrays or tensors (util.py:17-26)."""
    if isinstance(arr, list):
        print([(list(a.shape), f"{a.min():.2f}", f"{a.mean():.2f}", f"{a.max():.2f}") for a in arr])
    else:
        print(list(arr.shape), f"{arr.min():.2f}", f"{arr.mean():.2f}", f"{arr.max():.2f}")


def _no_plot(*args, **kwargs):
    return None


plot_signals = plot_spectra = plot_audio = plot_chroma_comparison = _no_plot
