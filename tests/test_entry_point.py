"""CPU: the drop-in entry points keep the reference's CLI and job-function surface (train_seg_semisup_mask_mt.py:16-42,
581-650; train_seg_semisup_ict.py:4-14, 508-577; train_seg_semisup_aug_mt.py:4-16, 515-577;
train_seg_semisup_vat_mt.py:4-17, 589-659; golden recorded from the unmodified reference by oracle/gen_golden.py)."""
import inspect
import json
import os

import click
import pytest

import train_seg_semisup_aug_mt
import train_seg_semisup_ict
import train_seg_semisup_mask_mt
import train_seg_semisup_vat_mt
from architectures import network_architectures

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'entry_point.json')))
# options this build adds (documented in the module docstring); everything else must be the reference's
EXTRA_OPTIONS = {'no_pretrained', 'ddp', 'synthetic_classes'}
EXTRA_CHOICES = {'dataset': {'synthetic', 'synthetic_u8'}}


SCRIPTS = {'train_seg_semisup_mask_mt': train_seg_semisup_mask_mt, 'train_seg_semisup_ict': train_seg_semisup_ict,
           'train_seg_semisup_aug_mt': train_seg_semisup_aug_mt, 'train_seg_semisup_vat_mt': train_seg_semisup_vat_mt}
entry = train_seg_semisup_mask_mt


def _ours(script):
    return {p.name: p for p in SCRIPTS[script].experiment.params}


@pytest.mark.parametrize('script', sorted(SCRIPTS))
def test_every_reference_option_is_kept_with_its_default_and_type(script):
    ours = _ours(script)
    for ref in GOLD['scripts'][script]['options']:
        assert ref['name'] in ours, 'missing option --{}'.format(ref['name'])
        p = ours[ref['name']]
        assert list(p.opts) == ref['opts']
        assert bool(getattr(p, 'is_flag', False)) == ref['is_flag'], ref['name']
        assert type(p.type).__name__ == ref['type'], ref['name']
        if not callable(p.default):
            assert p.default == ref['default'], ref['name']
        if ref['choices'] is not None:
            assert isinstance(p.type, click.Choice)
            mine = list(p.type.choices)
            assert mine[:len(ref['choices'])] == ref['choices'], ref['name']
            assert set(mine) - set(ref['choices']) <= EXTRA_CHOICES.get(ref['name'], set()), ref['name']


@pytest.mark.parametrize('script', sorted(SCRIPTS))
def test_only_documented_options_are_added(script):
    ref_names = {o['name'] for o in GOLD['scripts'][script]['options']}
    assert set(_ours(script)) - ref_names == EXTRA_OPTIONS
    for name in EXTRA_OPTIONS:            # additions must not change behaviour unless asked for
        p = _ours(script)[name]
        assert p.default in (False, 21)


@pytest.mark.parametrize('script', sorted(SCRIPTS))
def test_job_function_signature_and_submit(script):
    gold = GOLD['scripts'][script]
    job = getattr(SCRIPTS[script], script)
    params = list(inspect.signature(job).parameters)
    assert params[:len(gold['job_params'])] == gold['job_params']
    extra = params[len(gold['job_params']):]
    assert set(extra) == EXTRA_OPTIONS
    sig = inspect.signature(job)
    assert all(sig.parameters[k].default is not inspect.Parameter.empty for k in extra)      # reference callers still work
    assert gold['has_submit'] and callable(job.submit)


def test_sigmoid_rampup_known_answers():
    for e, want in enumerate(GOLD['rampup']):
        assert network_architectures.sigmoid_rampup(e, 10) == pytest.approx(want, rel=1e-12, abs=0.0)


def test_bad_mask_mode_is_rejected_before_any_gpu_work():
    # click validates the choice: the reference exits with usage error 2 as well
    from click.testing import CliRunner
    r = CliRunner().invoke(entry.experiment, ['--mask_mode', 'blend'])
    assert r.exit_code == 2


def test_real_datasets_fail_fast_with_a_usage_error_and_ignored_options_are_reported(tmp_path, monkeypatch):
    """the CLI default `pascal_aug` (and every real dataset) needs the reference's CPU data pipeline, which is not part
    of this build -> a click usage error before any device / process-group initialisation (this test runs without a GPU);
    data-pipeline options that cannot take effect on synthetic data are listed, not silently swallowed."""
    from click.testing import CliRunner
    import train_seg_semisup_mask_mt as entry
    from cutmix_semisup_seg_b200 import train_loop
    monkeypatch.chdir(tmp_path)
    r = CliRunner().invoke(entry.experiment, ['--num_epochs', '1'])
    assert r.exit_code == 2 and 'pascal_aug' in r.output and '--dataset synthetic' in r.output, r.output
    settings = dict(train_loop.DATA_PIPELINE_OPTIONS)
    assert train_loop.ignored_options(settings) == []
    settings.update(n_sup=372, aug_hflip=True, aug_rot_mag=10.0)
    assert sorted(train_loop.ignored_options(settings)) == ['aug_hflip', 'aug_rot_mag', 'n_sup']
    assert sorted(train_loop.ignored_options(settings, used=('aug_rot_mag',))) == ['aug_hflip', 'n_sup']
