"""Dynamic batch tokenizer for the asyncio front half (reference ``python/sglang/srt/managers/async_dynamic_batch_tokenizer.py``,
``--enable-dynamic-batch-tokenizer``): texts arriving within a short window are encoded with ONE batched tokenizer call on a worker thread —
the Hugging Face fast tokenizers parallelise a batch in Rust with the GIL released, and the event loop never blocks on a long prompt."""
from __future__ import annotations

import asyncio
from concurrent.futures import ThreadPoolExecutor
from typing import Any, List, Optional, Tuple


class AsyncDynamicBatchTokenizer:
    def __init__(self, tokenizer: Any, max_batch_size: int = 32, batch_wait_timeout_s: float = 0.002):
        self.tok, self.max_bs, self.wait = tokenizer, int(max_batch_size), float(batch_wait_timeout_s)
        self._q: Optional[asyncio.Queue] = None
        self._task: Optional[asyncio.Task] = None
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="sglb-tokenize")
        self.batches = 0
        self.items = 0

    def _encode_batch(self, texts: List[str]) -> List[List[int]]:
        if len(texts) > 1:
            try:
                out = self.tok(texts)            # fast tokenizers: one call, Rust-side parallelism
                ids = out["input_ids"] if isinstance(out, dict) or hasattr(out, "keys") else out
                if len(ids) == len(texts):
                    return [list(x) for x in ids]
            except Exception:
                pass
        return [list(self.tok.encode(t)) for t in texts]

    async def encode(self, text: str) -> List[int]:
        loop = asyncio.get_running_loop()
        if self._q is None or self._task is None or self._task.done() or self._task.get_loop() is not loop:
            self._q = asyncio.Queue()
            self._task = loop.create_task(self._run(self._q))
        fut: asyncio.Future = loop.create_future()
        await self._q.put((text, fut))
        return await fut

    async def _run(self, q: "asyncio.Queue[Tuple[str, asyncio.Future]]") -> None:
        loop = asyncio.get_running_loop()
        while True:
            batch = [await q.get()]
            deadline = loop.time() + self.wait
            while len(batch) < self.max_bs:
                left = deadline - loop.time()
                if left <= 0 and q.empty():
                    break
                try:
                    batch.append(q.get_nowait() if not q.empty() else await asyncio.wait_for(q.get(), max(left, 0)))
                except asyncio.TimeoutError:
                    break
            texts = [t for t, _ in batch]
            try:
                ids = await loop.run_in_executor(self._pool, self._encode_batch, texts)
                self.batches += 1
                self.items += len(texts)
                for (_, f), x in zip(batch, ids):
                    if not f.done():
                        f.set_result(x)
            except Exception as e:          # hand the failure to every waiter of this batch
                for _, f in batch:
                    if not f.done():
                        f.set_exception(e)
