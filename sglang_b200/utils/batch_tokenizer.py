"""Dynamic batching of tokenisation requests: concurrent ``encode`` calls issued within a short window are served by ONE batched call into
the (Rust-backed, GIL-releasing) fast tokenizer.

Parity: reference ``managers/async_dynamic_batch_tokenizer.py`` (``--enable-dynamic-batch-tokenizer``, batch size / timeout knobs)."""
from __future__ import annotations

import asyncio
from typing import Any, List, Tuple


class AsyncDynamicBatchTokenizer:
    def __init__(self, tokenizer: Any, max_batch_size: int = 32, batch_wait_timeout_s: float = 0.002):
        self.tk, self.max_bs, self.wait = tokenizer, max_batch_size, batch_wait_timeout_s
        self._q: "asyncio.Queue[Tuple[str, asyncio.Future]]" = None      # type: ignore[assignment]
        self._task = None
        self.batches: List[int] = []          # sizes of the executed batches (introspection / tests)

    def _ensure(self) -> None:
        if self._q is None:
            self._q = asyncio.Queue()
            self._task = asyncio.get_running_loop().create_task(self._loop())

    async def encode(self, text: str) -> List[int]:
        self._ensure()
        fut: asyncio.Future = asyncio.get_running_loop().create_future()
        await self._q.put((text, fut))
        return await fut

    async def _loop(self) -> None:
        loop = asyncio.get_running_loop()
        while True:
            item = await self._q.get()
            batch = [item]
            deadline = loop.time() + self.wait
            while len(batch) < self.max_bs:
                t = deadline - loop.time()
                if t <= 0:
                    break
                try:
                    batch.append(await asyncio.wait_for(self._q.get(), timeout=t))
                except asyncio.TimeoutError:
                    break
            texts = [b[0] for b in batch]
            try:
                ids = await loop.run_in_executor(None, self._encode_batch, texts)
                for (_, fut), row in zip(batch, ids):
                    if not fut.done():
                        fut.set_result(row)
            except Exception as e:      # noqa: BLE001
                for _, fut in batch:
                    if not fut.done():
                        fut.set_exception(e)
            self.batches.append(len(batch))

    def _encode_batch(self, texts: List[str]) -> List[List[int]]:
        if callable(self.tk) and hasattr(self.tk, "batch_encode_plus"):
            return [list(r) for r in self.tk(texts)["input_ids"]]
        return [list(self.tk.encode(t)) for t in texts]
