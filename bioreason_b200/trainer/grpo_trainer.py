"""B200-native `DNALLMGRPOTrainer`: the tensor math of bioreason/trainer/grpo_trainer.py on libbioreason_b200.

Kept from the reference: `RepeatRandomSampler` semantics (:72-119), the rollout with the hard-coded sampling config
(:384-391), EOS-inclusive completion mask (:605-609), ref log-probs with the frozen reference policy, old log-probs only
when num_iterations > 1 (:617-640), all-gather of rewards then group-normalised advantages with unbiased std and +1e-4
(:679-699), the clipped-ratio + beta*k3-KL loss with per-row masked mean (:786-812), the metric names (:703-716, :803, :812).
Re-designed: one fused lm_head+log-softmax+gather kernel instead of [B, L, V] logits; rollout on the paged-KV decode
engine with the G samples sharing one prefill; hand-written backward; one flat NCCL all-reduce of the LoRA+projector
gradients.  Not HF-Trainer based (accelerate/trl/peft are not installed in this image).
"""
from __future__ import annotations

import time
from collections import defaultdict
from typing import Any, Callable, Dict, List, Optional, Sized, Union

import torch
import torch.distributed as dist
from torch.utils.data import Sampler

from .. import dp, ops, training
from . import rewards as rw
from .grpo_config import DNALLMGRPOConfig


class RepeatRandomSampler(Sampler):
    """grpo_trainer.py:72-119 -- each index repeated `mini_repeat_count` times, chunks of `batch_size` unique indices,
    the whole chunk repeated `repeat_count` times; same seed on every rank."""

    def __init__(self, data_source: Sized, mini_repeat_count: int, batch_size: int = 1, repeat_count: int = 1, seed: Optional[int] = None):
        self.data_source, self.mini_repeat_count, self.batch_size, self.repeat_count = data_source, mini_repeat_count, batch_size, repeat_count
        self.num_samples = len(data_source)
        self.seed = seed
        self.generator = torch.Generator()
        if seed is not None:
            self.generator.manual_seed(seed)

    def __iter__(self):
        indexes = torch.randperm(self.num_samples, generator=self.generator).tolist()
        indexes = [indexes[i:i + self.batch_size] for i in range(0, len(indexes), self.batch_size)]
        indexes = [chunk for chunk in indexes if len(chunk) == self.batch_size]
        for chunk in indexes:
            for _ in range(self.repeat_count):
                for index in chunk:
                    for _ in range(self.mini_repeat_count):
                        yield index

    def __len__(self) -> int:
        return self.num_samples * self.mini_repeat_count * self.repeat_count


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


class TrainerState:
    """The fields of transformers.TrainerState that callbacks on this path read (reason.py:46-81 uses global_step)."""

    def __init__(self):
        self.global_step, self.epoch, self.max_steps = 0, 0.0, 0
        self.log_history: List[Dict[str, float]] = []
        self.is_world_process_zero = _world()[0] == 0
        self.is_local_process_zero = self.is_world_process_zero


class TrainerControl:
    def __init__(self):
        self.should_save = self.should_log = self.should_training_stop = self.should_evaluate = self.should_epoch_stop = False


class CallbackHandler:
    """Duck-typed transformers.TrainerCallback dispatch: event(args, state, control, model=, processing_class=, optimizer=, ...);
    a callback may return a (modified) control object."""

    def __init__(self, callbacks, trainer):
        self.callbacks, self.trainer = list(callbacks or []), trainer

    def fire(self, event: str, **extra):
        tr = self.trainer
        for cb in self.callbacks:
            fn = getattr(cb, event, None)
            if fn is None:
                continue
            out = fn(tr.args, tr.state, tr.control, model=tr.model, processing_class=tr.processing_class, tokenizer=tr.processing_class,
                     optimizer=tr.optimizer, train_dataloader=None, eval_dataloader=None, **extra)
            if out is not None:
                tr.control = out
        return tr.control


class DNALLMGRPOTrainer:
    def __init__(self, model, reward_funcs: Union[Callable, List[Callable]], args: Optional[DNALLMGRPOConfig] = None, dna_module=None,
                 train_dataset=None, eval_dataset=None, processing_class=None, reward_processing_classes=None, callbacks=None,
                 optimizers=(None, None), peft_config=None, freeze_dna_modules: bool = False, attn_implementation: str = "flash_attention_2",
                 torch_dtype: str = "bfloat16", **kwargs):
        assert not isinstance(model, str), "model must be a DNALLMModel instance"             # grpo_trainer.py:241
        self.model, self.args = model, args or DNALLMGRPOConfig()
        a = self.args
        self.reward_funcs = list(reward_funcs) if isinstance(reward_funcs, (list, tuple)) else [reward_funcs]
        self.dna_module, self.processing_class = dna_module, processing_class
        self.train_dataset, self.eval_dataset = train_dataset, eval_dataset
        self.num_generations, self.max_completion_length = a.num_generations, a.max_completion_length
        self.beta, self.num_iterations = a.beta, a.num_iterations
        self.epsilon_low = a.epsilon
        self.epsilon_high = a.epsilon_high if a.epsilon_high is not None else a.epsilon
        rank, world = _world()
        global_bs = a.per_device_train_batch_size * world
        possible = [n for n in range(2, global_bs + 1) if global_bs % n == 0]
        if self.num_generations not in possible:                                            # grpo_trainer.py:428-436
            raise ValueError(f"The global train batch size ({world} x {a.per_device_train_batch_size}) must be evenly divisible by the "
                             f"number of generations per prompt ({self.num_generations}). Given the current train batch size, the valid "
                             f"values for the number of generations are: {possible}.")
        if model._lora is None:
            model.enable_lora(r=a.lora_r, alpha=a.lora_alpha, seed=a.seed)
        model.sync_adapters(rollout=True)
        self.eos_token_id = getattr(processing_class, "eos_token_id", None) if processing_class is not None else None
        if self.eos_token_id is None:
            self.eos_token_id = model.text_config.eos_token_id
        self.pad_token_id = getattr(processing_class, "pad_token_id", None) if processing_class is not None else None
        if self.pad_token_id is None:
            self.pad_token_id = model.text_config.pad_token_id
        # hard-coded exactly like grpo_trainer.py:384-391 (args.temperature/top_p/top_k are NOT consulted there either)
        self.generation_kwargs = dict(max_new_tokens=self.max_completion_length, do_sample=True, temperature=0.6, top_p=0.95, top_k=20,
                                      pad_token_id=self.pad_token_id, eos_token_id=None if a.suppress_eos else self.eos_token_id)
        opt = optimizers[0]
        if opt is None:
            opt = torch.optim.AdamW(model.trainable_parameters(), lr=a.learning_rate, betas=(a.adam_beta1, a.adam_beta2), eps=a.adam_epsilon,
                                    weight_decay=a.weight_decay, fused=True)
        self.optimizer = opt
        self._metrics = defaultdict(list)
        self._buffered_inputs = [None] * a.gradient_accumulation_steps
        self._step = 0
        self.state, self.control = TrainerState(), TrainerControl()
        self.state.max_steps = a.max_steps
        self.callback_handler = CallbackHandler(callbacks, self)
        self.reward_d2h_bytes = 0          # bytes of completion ids copied to the host for text reward functions (last step)
        # per-rank distinct sampling stream (set_seed(seed, device_specific=True), grpo_trainer.py:451)
        self._gen = torch.Generator(device="cuda")
        self._gen.manual_seed(a.seed + rank)
        self.timings = defaultdict(float)
        self._ev = []                      # (phase, start_event, end_event): GPU-side phase times, read by gpu_phase_ms()

    @property
    def global_step(self):
        return self.state.global_step

    @global_step.setter
    def global_step(self, v):
        self.state.global_step = v

    def add_callback(self, cb):
        self.callback_handler.callbacks.append(cb)

    def _mark(self, phase):
        """Context manager: CUDA-event bracket of a phase on the current stream (no host sync)."""
        tr = self

        class _M:
            def __enter__(self_m):
                self_m.e0 = torch.cuda.Event(enable_timing=True); self_m.e1 = torch.cuda.Event(enable_timing=True)
                self_m.e0.record()

            def __exit__(self_m, *a):
                self_m.e1.record()
                tr._ev.append((phase, self_m.e0, self_m.e1))
        return _M()

    def gpu_phase_ms(self, reset=True):
        torch.cuda.synchronize()
        out = defaultdict(float)
        for ph, e0, e1 in self._ev:
            out[ph] += e0.elapsed_time(e1)
        if reset:
            self._ev = []
        return dict(out)

    # ------------------------------------------------------------------ data
    def _get_train_sampler(self):                                                          # grpo_trainer.py:883-897
        _, world = _world()
        a = self.args
        eff = a.per_device_train_batch_size * world * a.gradient_accumulation_steps
        return RepeatRandomSampler(self.train_dataset, self.num_generations, eff // self.num_generations, self.num_iterations, a.seed)

    def _prepare_prompt_inputs(self, inputs) -> Dict[str, Any]:
        """Pre-tokenised batches pass through; raw examples go through dna_module + processor like :538-567.  Returns the model
        inputs plus `_examples` (the raw example dicts, for the reward columns) and `_prompts`."""
        if isinstance(inputs, dict) and "input_ids" in inputs:
            out = dict(inputs)
            out.setdefault("_examples", inputs.get("examples"))
            out.setdefault("_prompts", inputs.get("prompts"))
            return out
        if self.dna_module is None or self.processing_class is None:
            raise ValueError("raw examples need dna_module and processing_class (no tokenizer files exist offline); pass a tokenised batch")
        prompts_text = self.dna_module.prepare_prompt(self.processing_class, inputs)
        dnas = [x["dna_sequences"] for x in inputs]
        out = dict(self.dna_module.prepare_model_inputs(self.processing_class, self.model, prompts_text, dnas, return_tensors="pt", padding=True,
                                                        padding_side="left", add_special_tokens=False))
        out["_examples"] = list(inputs)
        out["_prompts"] = [x["prompt"] for x in inputs]                                     # grpo_trainer.py:537
        return out

    # ------------------------------------------------------------------ log-probs
    def _get_per_token_logps(self, model, input_ids, attention_mask, keep_last=None, lora="policy", **mm):
        """grpo_trainer.py:510-520 (+ the [:, P-1:] slice of :779 when keep_last is given), no-grad version."""
        n = input_ids.shape[1] - 1 if keep_last is None else keep_last
        with torch.no_grad():
            lp, _ = training.policy_forward(model, input_ids, attention_mask, mm.get("dna_tokenized"), mm.get("batch_idx_map"), n,
                                            save=False, lora=lora)
        return lp

    # ------------------------------------------------------------------ rollout + scoring
    @torch.no_grad()
    def _generate_and_score_completions(self, inputs, model, uniforms=None, rewards_per_func=None) -> Dict[str, Any]:
        t0 = time.perf_counter()
        pi = self._prepare_prompt_inputs(inputs)
        dev = model._dec.embed.device
        prompt_ids, prompt_mask = pi["input_ids"].to(dev), pi["attention_mask"].to(dev)
        dna = pi.get("dna_tokenized")
        if dna is not None:                                                # device-resident once: the three passes of the step reuse the tensors
            dna = {k: dna[k].to(dev) for k in ("input_ids", "attention_mask")}
        mm = dict(dna_tokenized=dna, batch_idx_map=pi.get("batch_idx_map"))
        B, P = prompt_ids.shape
        C = self.max_completion_length
        if uniforms is None:
            uniforms = torch.rand(C, B, device=dev, generator=self._gen)
        with self._mark("rollout"):
            completion_ids = model.generate(prompt_ids, prompt_mask, mm["dna_tokenized"], mm["batch_idx_map"], uniforms=uniforms, **self.generation_kwargs)
        self.timings["rollout"] += time.perf_counter() - t0
        completion_mask = ops.eos_mask(completion_ids, self.eos_token_id if not self.args.suppress_eos else -1)      # :605-609
        # text reward functions need the ids on the host: start the copy now (side stream, pinned), wait for it only after the
        # reference-policy forward has been queued -> decode + CPU rewards overlap that forward
        need_text = rewards_per_func is None and any(not rw.wants_token_protocol(f) for f in self.reward_funcs)
        host_copy = rw.AsyncHostCopy(completion_ids) if need_text else None
        self.reward_d2h_bytes = host_copy.nbytes if host_copy is not None else 0
        ids = torch.cat([prompt_ids, completion_ids], dim=1)
        attention_mask = torch.cat([prompt_mask, completion_mask.to(prompt_mask.dtype)], dim=1)                       # :612
        Cc = completion_ids.shape[1]
        with self._mark("ref_logps"):
            old_lp = self._get_per_token_logps(model, ids, attention_mask, keep_last=Cc, **mm) if self.num_iterations > 1 else None
            ref_lp = self._get_per_token_logps(model, ids, attention_mask, keep_last=Cc, lora=None, **mm) if self.beta != 0.0 else None
        # rewards: the reference protocol f(prompts=, completions=, **columns) on decoded text (:640-676); functions that name a
        # `completion_ids` parameter get device tensors instead (trainer/rewards.py)
        if rewards_per_func is None:
            t_r = time.perf_counter()
            rewards_per_func = rw.score(self.reward_funcs, examples=pi.get("_examples"), prompts=pi.get("_prompts"), completion_ids=completion_ids,
                                        completion_mask=completion_mask, prompt_ids=prompt_ids, processing_class=self.processing_class,
                                        host_copy=host_copy, extra_columns=pi.get("reward_kwargs"))
            self.timings["reward_host"] += time.perf_counter() - t_r
        rewards_all = dp.gather_rewards(rewards_per_func)                                                              # C1, :679
        adv_all, gmean, gstd = ops.grpo_advantages(rewards_all, self.num_generations, return_stats=True)               # :682-692
        advantages = dp.local_slice(adv_all, B)                                                                        # :695-699
        self._metrics["completion_length"].append(completion_mask.sum(1).float().mean())
        self._metrics["reward"].append(rewards_all.sum(1).mean())
        self._metrics["reward_std"].append(gstd.mean())
        for i, f in enumerate(self.reward_funcs):
            self._metrics[f"rewards/{getattr(f, '__name__', 'reward_' + str(i))}"].append(rewards_all[:, i].mean())
        self.timings["score"] += time.perf_counter() - t0
        return dict(prompt_ids=prompt_ids, prompt_mask=prompt_mask, completion_ids=completion_ids, completion_mask=completion_mask,
                    old_per_token_logps=old_lp, ref_per_token_logps=ref_lp, advantages=advantages, multimodal_inputs=mm)

    # ------------------------------------------------------------------ loss (+ backward through the kernels)
    def compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None, backward: bool = True):
        if return_outputs:
            raise ValueError("The GRPOTrainer does not support returning outputs")          # grpo_trainer.py:752-753
        if self.global_step % self.num_iterations == 0 or self._buffered_inputs[self._step % self.args.gradient_accumulation_steps] is None:
            if "completion_ids" not in inputs:
                inputs = self._generate_and_score_completions(inputs, model)
            self._buffered_inputs[self._step % self.args.gradient_accumulation_steps] = inputs
        else:
            inputs = self._buffered_inputs[self._step % self.args.gradient_accumulation_steps]
        self._step += 1
        prompt_ids, prompt_mask = inputs["prompt_ids"], inputs["prompt_mask"]
        completion_ids, completion_mask = inputs["completion_ids"], inputs["completion_mask"]
        mm = inputs["multimodal_inputs"]
        ids = torch.cat([prompt_ids, completion_ids], dim=1)
        mask = torch.cat([prompt_mask, completion_mask.to(prompt_mask.dtype)], dim=1)
        B, C = completion_ids.shape
        adv, old, ref = inputs["advantages"], inputs["old_per_token_logps"], inputs["ref_per_token_logps"]
        mr = self.args.micro_rows or B
        ga = self.args.gradient_accumulation_steps
        loss_acc = torch.zeros(3, device=ids.device)
        for lo in range(0, B, mr):
            hi = min(B, lo + mr)
            sl = slice(lo, hi)
            mm_c = _slice_mm(mm, lo, hi)
            t0 = time.perf_counter()
            with self._mark("policy_fwd"):
                lp, ctx = training.policy_forward(model, ids[sl], mask[sl], mm_c["dna_tokenized"], mm_c["batch_idx_map"], C, save=backward)
            out3, dlp = ops.grpo_loss_raw(lp, old[sl] if old is not None else None, ref[sl] if ref is not None else None, adv[sl],
                                          completion_mask[sl], self.beta, self.epsilon_low, self.epsilon_high, want_grad=backward)
            w = (hi - lo) / B
            loss_acc += out3 * w                                            # row-mean of row-means is separable over row chunks
            self.timings["policy_fwd"] += time.perf_counter() - t0
            if backward:
                t0 = time.perf_counter()
                # the last chunk of the last accumulation micro-step completes the gradients: all-reduce each layer's slice as the
                # backward leaves it (C2 overlapped with the remaining backward)
                hook = None
                if hi == B and self._step % ga == 0 and _world()[1] > 1 and getattr(model, "_lora", None) is not None:
                    self._reducer = dp.OverlappedGradReduce(model._lora.flat_grad)
                    hook = lambda li, _m=model: self._reducer.reduce_slice(*_m._lora.layer_slice(li))
                with self._mark("policy_bwd"):
                    training.policy_backward(model, ctx, dlp * (w / ga), on_layer_done=hook)
                self.timings["policy_bwd"] += time.perf_counter() - t0
        # clip_ratio is a ratio of sums; with row chunks it is weighted by rows (exact when chunks have equal mask counts)
        if self.beta > 0:
            self._metrics["kl"].append(loss_acc[1])
        self._metrics["clip_ratio"].append(loss_acc[2])
        return loss_acc[0]

    # ------------------------------------------------------------------ one optimizer step
    def training_step(self, inputs) -> torch.Tensor:
        model = self.model
        if self._step % self.args.gradient_accumulation_steps == 0:
            model.zero_grad_buffers()
        loss = self.compute_loss(model, inputs)
        if self._step % self.args.gradient_accumulation_steps == 0:
            self._optimizer_step()
        return loss

    def _optimizer_step(self):
        model = self.model
        t0 = time.perf_counter()
        with self._mark("grad_allreduce"):
            red = getattr(self, "_reducer", None)
            if red is not None:                                            # slices already in flight under the backward; wait + the rest
                red.finish([model._proj_grad_w, model._proj_grad_b])
                self._reducer = None
            else:
                dp.allreduce_mean_([model._lora.flat_grad, model._proj_grad_w, model._proj_grad_b])     # C2: sum then / world (DDP average)
        with self._mark("optimizer"):
            model.attach_grads()
            if self.args.max_grad_norm and self.args.max_grad_norm > 0:
                torch.nn.utils.clip_grad_norm_(model.trainable_parameters(), self.args.max_grad_norm, foreach=True)
            self.optimizer.step()
        with self._mark("adapter_sync"):
            model.sync_adapters(rollout=True)
        self.global_step += 1
        self.timings["optimizer"] += time.perf_counter() - t0
        a = self.args
        self.control = self.callback_handler.fire("on_step_end")
        if getattr(a, "logging_steps", 0) and self.callback_handler.callbacks and self.global_step % max(1, int(a.logging_steps)) == 0:
            logs = self.log_metrics()
            self.state.log_history.append(dict(logs, step=self.global_step))
            self.control = self.callback_handler.fire("on_log", logs=logs)
        if getattr(a, "save_steps", 0) and self.global_step % int(a.save_steps) == 0:
            self.control.should_save = True
        if self.control.should_save:
            self.control = self.callback_handler.fire("on_save")               # reason.py:46-81 SaveWithPyTorchCallback hooks here
            self.control.should_save = False

    def train(self, batches=None, max_steps: Optional[int] = None):
        """Iterate tokenised batches (or the dataset through RepeatRandomSampler) for max_steps optimizer steps."""
        a = self.args
        steps = max_steps if max_steps is not None else (a.max_steps if a.max_steps > 0 else None)
        if batches is None:
            sampler = list(iter(self._get_train_sampler()))
            batches = ([self.train_dataset[i] for i in idx] for idx in dp.rank_batches(sampler, a.per_device_train_batch_size))
        out = []
        self.control = self.callback_handler.fire("on_train_begin")
        for b in batches:
            self.control = self.callback_handler.fire("on_step_begin")
            out.append(self.training_step(b))
            if (steps is not None and self.global_step >= steps) or self.control.should_training_stop:
                break
        self.control = self.callback_handler.fire("on_train_end")
        return out

    def log_metrics(self) -> Dict[str, float]:
        m = {k: float(torch.stack([torch.as_tensor(x, dtype=torch.float32, device="cuda") for x in v]).mean()) for k, v in self._metrics.items()}
        self._metrics.clear()
        return m


def _slice_mm(mm, lo, hi):
    """Row-chunk the multimodal inputs (dna rows follow batch_idx_map)."""
    if mm.get("dna_tokenized") is None or not mm.get("batch_idx_map"):
        return dict(dna_tokenized=None, batch_idx_map=[])
    idx = [i for i, b in enumerate(mm["batch_idx_map"]) if lo <= b < hi]
    it = torch.tensor(idx, device=mm["dna_tokenized"]["input_ids"].device)
    return dict(dna_tokenized={k: v[it] for k, v in mm["dna_tokenized"].items() if k in ("input_ids", "attention_mask")},
                batch_idx_map=[mm["batch_idx_map"][i] - lo for i in idx])
