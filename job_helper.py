"""Minimal stand-in for the reference's job_helper (job_helper.py:14-146): `@job(name)` adds `.submit(**kw)`,
which creates results/<job>/, tees stdout/stderr into log_<desc>.txt and refuses to re-run a finished job."""
import os
import sys


class LogAlreadyExistsError(Exception):
    pass


class Logger(object):
    """Tee: everything written goes to the log file and to the wrapped stream."""

    def __init__(self, path, stream):
        self.path, self.stream = path, stream

    def write(self, x):
        with open(self.path, 'a+') as f:
            f.write(x)
        self.stream.write(x)

    def flush(self):
        self.stream.flush()


class SubmitConfig(object):
    def __init__(self, job_name, job_desc, enumerate_job_names):
        res_dir = os.path.join('results', job_name)
        os.makedirs(res_dir, exist_ok=True)
        if job_desc == 'none':
            self.log_path = self.job_out_dir = None
        else:
            if enumerate_job_names:
                taken = [-1]
                for name in os.listdir(res_dir):
                    stem = name[4:] if name.startswith('log_') else name
                    digits = stem.split('_')[0]
                    if digits.isdigit():
                        taken.append(int(digits))
                index = max(taken) + 1
                self.log_path = os.path.join(res_dir, 'log_{:04d}_{}.txt'.format(index, job_desc))
                self.job_out_dir = os.path.join(res_dir, '{:04d}_{}'.format(index, job_desc))
            else:
                self.log_path = os.path.join(res_dir, 'log_{}.txt'.format(job_desc))
                self.job_out_dir = os.path.join(res_dir, job_desc)
                if os.path.exists(self.log_path) or os.path.exists(self.job_out_dir):
                    raise LogAlreadyExistsError
        self._run_dir = None
        self._streams = None

    @property
    def run_dir(self):
        if self._run_dir is None and self.job_out_dir is not None:
            self._run_dir = self.job_out_dir
            os.makedirs(self._run_dir, exist_ok=True)
        return self._run_dir

    def connect_streams(self):
        if self.log_path is not None:
            self._streams = (sys.stdout, sys.stderr)
            sys.stdout = Logger(self.log_path, sys.stdout)
            sys.stderr = Logger(self.log_path, sys.stderr)

    def disconnect_streams(self):
        if self._streams is not None:
            sys.stdout, sys.stderr = self._streams
            self._streams = None


def job(job_name, enumerate_job_names=True):
    def decorate(job_fn):
        def run_job(**kwargs):
            name = kwargs.pop('job_name', None) or job_name
            quota_group = kwargs.pop('quota_group', None)
            if quota_group:
                raise ValueError('quota_group not supported when dnnlib is not available')
            desc = kwargs.pop('job_desc', None) or name
            try:
                cfg = SubmitConfig(name, desc, enumerate_job_names)
            except LogAlreadyExistsError:
                print('Job {}:{} already executed; skipping'.format(name, desc))
                return
            print('[NO dnnlib] logging to {}'.format(cfg.log_path))
            cfg.connect_streams()
            try:
                job_fn(cfg, **kwargs)
            finally:
                cfg.disconnect_streams()
        job_fn.submit = run_job
        return job_fn
    return decorate
