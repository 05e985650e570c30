"""CPU: host-side mirror of the reference module API (names, ctor behaviour, state_dict layout)."""
import pytest
import torch

from tests import helpers as H


def test_state_dict_layout_matches_reference():
    m = H.seeded_model(128, 3)
    sd = m.state_dict()
    assert len(sd) == 84 and sum(p.numel() for p in m.parameters()) == 7312707
    assert tuple(sd["decoder.net.6.weight"].shape) == (128, 3, 5, 5)
    assert tuple(sd["hyper_decoder.net.2.weight"].shape) == (128, 192, 5, 5)
    assert tuple(sd["hyper_decoder.net.4.weight"].shape) == (256, 192, 3, 3)
    assert tuple(sd["context_model.masked.mask"].shape) == (256, 128, 5, 5)
    assert tuple(sd["entropy_parameters.net.4.weight"].shape) == (1152, 640, 1, 1)
    assert tuple(sd["factorized_entropy_model.matrices.1"].shape) == (128, 3, 3)
    for k in ("beta", "gamma", "beta_reparam.pedestal", "beta_reparam.lower_bound.bound",
              "gamma_reparam.pedestal", "gamma_reparam.lower_bound.bound"):
        assert f"encoder.net.1.{k}" in sd
    assert H.state_digest(sd) == str(H.load_golden("c2_k3_128x192_plain")["state_digest"])


def test_k1_variant_and_attributes():
    from neural_image_compression_b200.EntropyModels import GaussianConditional, GaussianMixtureConditional
    m1 = H.seeded_model(128, 1)
    assert sum(p.numel() for p in m1.parameters()) == 6738371
    assert isinstance(m1.conditional, GaussianConditional) and m1.distribution == "Mean-Scale Gaussian"
    m3 = H.seeded_model(128, 3)
    assert isinstance(m3.conditional, GaussianMixtureConditional) and m3.distribution == "Mixture of Gaussians"
    for attr in ("encoder", "decoder", "hyper_encoder", "hyper_decoder", "factorized_entropy_model", "context_model",
                 "entropy_parameters", "conditional", "M", "K", "H"):
        assert hasattr(m3, attr)
    assert m3.factorized_entropy_model.likelihood_bound == 1e-9


def test_constructor_errors_match_reference():
    from neural_image_compression_b200.Models import JointAutoregressiveHierarchical as J
    from neural_image_compression_b200.ParametersModels import EntropyParameters
    for bad in (0, -1, 1.5, "a"):
        with pytest.raises(ValueError):
            J(bad, K=1)
        with pytest.raises(ValueError):
            J(128, K=bad)
        with pytest.raises(ValueError):
            EntropyParameters(128, 128, K=bad)


def test_channel_cdf_pmf_diagnostics_run_on_cpu():
    m = H.seeded_model(128, 1)
    x = torch.arange(-5, 6, dtype=torch.float32)
    cdf = m.factorized_entropy_model.channel_cdf(3, x)
    pmf = m.factorized_entropy_model.channel_pmf(3, x)
    assert cdf.shape == x.shape and torch.all(cdf[1:] >= cdf[:-1]) and torch.all(pmf > 0)
    # against the oracle's cumulative
    from oracle import forward as O
    ref = torch.sigmoid(O.factorized_logits(m.state_dict(), x.reshape(1, 1, -1).expand(128, 1, -1)))[3, 0]
    assert torch.allclose(cdf, ref, atol=1e-6)
