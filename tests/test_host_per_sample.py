"""Per-sample timesteps in the no-grad forward (per_sample.py; reference: sd3_5.py:394, nft.py:366-374): grouping and reassembly, on CPU
with a stand-in for the engine step, so that the row -> timestep association is pinned without a GPU."""
import torch

from flow_factory_b200.per_sample import forward_grouped, split_by_timestep
from flow_factory_b200.scheduler import SDESchedulerOutput


def test_uniform_batches_are_not_split():
    assert split_by_timestep(torch.tensor(500.0), None, 4) is None
    assert split_by_timestep(torch.tensor([500.0]), torch.tensor([400.0]), 4) is None
    assert split_by_timestep(torch.full((4,), 321.5), torch.full((4,), 300.0), 4) is None
    assert split_by_timestep(987.0, 900.0, 3) is None


def test_groups_cover_every_row_once_in_first_occurrence_order():
    t = torch.tensor([700.0, 100.0, 700.0, 350.0, 100.0])
    g = split_by_timestep(t, None, 5)
    assert [r.tolist() for r, _, _ in g] == [[0, 2], [1, 4], [3]]
    assert [float(tv) for _, tv, _ in g] == [700.0, 100.0, 350.0]
    assert all(tn is None for _, _, tn in g)
    # t shared, t_next per sample
    g = split_by_timestep(torch.tensor(700.0), torch.tensor([650.0, 600.0, 650.0]), 3)
    assert [r.tolist() for r, _, _ in g] == [[0, 2], [1]] and [float(tn) for _, _, tn in g] == [650.0, 600.0]


def test_wrong_length_is_rejected():
    try:
        split_by_timestep(torch.tensor([1.0, 2.0, 3.0]), None, 4)
    except ValueError:
        return
    raise AssertionError("a (3,) timestep tensor for a batch of 4 must be rejected")


def test_forward_grouped_evaluates_each_row_at_its_own_timestep():
    B = 6
    t = torch.tensor([900.0, 10.0, 900.0, 450.0, 10.0, 77.0])
    tn = t - 5.0
    latents = torch.arange(B, dtype=torch.float32).reshape(B, 1, 1, 1).expand(B, 2, 3, 3).contiguous()
    embeds = torch.arange(B, dtype=torch.float32).reshape(B, 1, 1) * 10
    calls = []

    def fake_forward(t, t_next, latents, prompt_embeds, guidance_scale, tags):
        calls.append((float(t), float(t_next), latents.shape[0], list(tags)))
        assert t.numel() == 1 and t_next.numel() == 1 and guidance_scale == 4.5
        return SDESchedulerOutput(next_latents=latents + t + 0.5 * t_next, log_prob=latents.flatten(1).mean(1) + prompt_embeds.flatten(1).mean(1) + t)

    out = forward_grouped(fake_forward, split_by_timestep(t, tn, B), B,
                          dict(latents=latents, prompt_embeds=embeds, guidance_scale=4.5, tags=[f"s{i}" for i in range(B)]),
                          batched=("latents", "prompt_embeds", "tags"), make_output=SDESchedulerOutput.from_dict)
    assert len(calls) == 4 and sorted(c[2] for c in calls) == [1, 1, 2, 2]
    assert calls[0][3] == ["s0", "s2"] and calls[1][3] == ["s1", "s4"]
    want = latents + t.view(B, 1, 1, 1) + 0.5 * tn.view(B, 1, 1, 1)
    assert torch.equal(out.next_latents, want)
    assert torch.equal(out.log_prob, torch.arange(B, dtype=torch.float32) * 11 + t)
    assert out.noise_pred is None
