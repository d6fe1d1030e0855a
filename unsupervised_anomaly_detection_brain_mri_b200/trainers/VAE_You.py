"""VAE_You trainer (mirror of reference trainers/VAE_You.py): the VAE objective (:47-52) plus iterative MAP restoration
with a total-variation prior at test time (:53-54, 125-147, 149-173).

The reference restores by 150 x sess.run per call, shipping the image to the device and the gradient back every
iteration; here the stack of slices is uploaded once and all iterations run as CUDA-graph replays of
``ConvAutoencoderEngine.restore_step`` (forward, dL/dxhat seed incl. the TV term, dgrad-only chain to the input, update)."""
import numpy as np

from .AEMODEL import AEMODEL, Phase, indicate_early_stopping, update_log_dicts  # noqa: F401


class VAE_You(AEMODEL):
    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('VAE_You')
            self.restore_lr = 1e-3
            self.restore_steps = 150
            self.tv_lambda = 1.8

    def __init__(self, sess, config=None, network=None):
        super().__init__(sess, config, network)
        self.restore_lr = self.config.restore_lr
        self.restore_steps = self.config.restore_steps
        self.tv_lambda_value = self.config.tv_lambda
        self.z_mu = self.outputs['z_mu']
        self.z_sigma = self.outputs['z_sigma']

    def train(self, dataset):
        super().train(dataset)
        if self.tv_lambda_value == -1 and self.restore_steps > 0:       # VAE_You.py:88-93
            print('Determining best lambda')
            self.determine_best_lambda(dataset)

    def _restore_stack(self, x, tv_lambda, dropout):
        """Restores a stack [N,H,W,C] in chunks of the evaluation batch size; returns the restored stack."""
        N = x.shape[0]
        chunk = min(N, int(getattr(self.config, 'evalBatchsize', 128)))
        eng = self._eval_engine(chunk)
        out = np.empty_like(x)
        for i in range(0, N, chunk):
            xb = x[i:i + chunk]
            n = xb.shape[0]
            if n < chunk:
                xb = np.concatenate([xb, np.zeros((chunk - n,) + xb.shape[1:], np.float32)], 0)
            eng.set_inputs(xb)
            eng.restore(int(self.restore_steps), float(self.restore_lr), float(tv_lambda), dropout=bool(dropout),
                        dropout_rate=float(self.config.dropout_rate),
                        use_graph=bool(getattr(self.config, 'use_cuda_graph', True)))
            out[i:i + n] = eng.br[0].x.cpu().numpy()[:n]
        return out

    def reconstruct(self, x, dropout=False):
        """VAE_You.py:125-147: 'reconstruction' is the RESTORED input, not x_hat."""
        if x.ndim < 4:
            x = np.expand_dims(x, 0)
        x = np.ascontiguousarray(x, np.float32)
        restored = self._restore_stack(x, self.tv_lambda_value, dropout)
        results = {'reconstruction': restored}
        results['l1err'] = np.sum(np.abs(x - restored))
        results['l2err'] = np.sum(np.sqrt((x - restored) ** 2))
        return results

    def determine_best_lambda(self, dataset):
        """VAE_You.py:149-173: grid over lambda in {0, 0.1, ..., 1.9} on 20 % of the validation batches."""
        lambdas = np.arange(20) / 10.0
        mean_errors = []
        nb = int(dataset.num_batches(self.config.batchsize, set=Phase.VAL.value) * 0.2)
        for tv_lambda in lambdas:
            errors = []
            for _ in range(nb):
                batch, _, _ = dataset.next_batch(self.config.batchsize, set=Phase.VAL.value)
                batch = np.ascontiguousarray(batch, np.float32)
                restored = self._restore_stack(batch, tv_lambda, False)
                errors.append(np.sum(np.abs(batch - restored)))
            mean_error = np.mean(errors) if errors else np.nan
            mean_errors.append(mean_error)
            print(f'mean_error for lambda {tv_lambda}: {mean_error}')
        self.tv_lambda_value = lambdas[mean_errors.index(min(mean_errors))]
        print(f'Best lambda: {self.tv_lambda_value}')
