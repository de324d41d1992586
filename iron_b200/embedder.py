"""NeRF positional encoding [x, sin(2^k x), cos(2^k x)]_{k<L} with the reference's interface
(models/embedder.py:6-54).  The networks of this package encode inside their CUDA kernels; this torch
version exists for callers that use `get_embedder` directly and to report the encoded width."""
import torch


class Embedder:
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs["input_dims"]
        n = kwargs["num_freqs"]
        max_freq = kwargs["max_freq_log2"]
        if kwargs.get("log_sampling", True):
            self.freq_bands = 2.0 ** torch.linspace(0.0, max_freq, n)
        else:
            self.freq_bands = torch.linspace(2.0 ** 0.0, 2.0 ** max_freq, n)
        self.periodic_fns = kwargs.get("periodic_fns", [torch.sin, torch.cos])
        self.include_input = kwargs.get("include_input", True)
        self.out_dim = (d if self.include_input else 0) + d * n * len(self.periodic_fns)

    def embed(self, inputs):
        parts = [inputs] if self.include_input else []
        for freq in self.freq_bands.tolist():
            for fn in self.periodic_fns:
                parts.append(fn(inputs * freq))
        return torch.cat(parts, -1)


def get_embedder(multires, input_dims=3):
    eo = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, eo=eo):
        return eo.embed(x)

    return embed, eo.out_dim
