// Host side of the PeerDAS / EIP-7594 cell calls (included by lwkzg.cu inside its anonymous namespace; the kernels
// are in cells.cu / cells_verify.cu, the C ABI wrappers in cells_api.inl).  Like the rest of the runtime this only
// moves bytes and launches kernels.

struct CellCtx {
  int c = 0, nwin = 0;            // window of the FK20 digit table
  void* d_tw = nullptr;           // w8192^k, k < 8192 (Montgomery)
  void* d_naf = nullptr;          // signed-digit recoding of the 128th roots of unity (G1 FFT twiddles)
  void* d_table = nullptr;        // GLV digit table of the 8192 FK20 points X[j * 64 + b]
  void* d_prep64 = nullptr;       // prepared Miller-loop lines of g2[64] = [tau^64]G2
  void* d_fk20 = nullptr;         // the 8192 FK20 points themselves (affine Montgomery, 768 KB): test hook
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr, ev_v[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // ev_v: fork / join points of a batched cell verification
  DevBuf blobs, coef, scalars, pts, cells, proofs, status, msm_scratch;
  // verification / recovery workspace
  DevBuf v_commit, v_cidx, v_cellidx, v_cells, v_proofs, v_evals, v_status, v_pts, v_wcoef, v_rpow, v_scal, v_r, v_out, v_ok, v_msm_a, v_msm_b;
  void* h_stage = nullptr;
  size_t h_stage_cap = 0;
};

// temporaries of the set-up / test-hook paths: freed on every exit, also the error ones
struct ScratchMem {
  std::vector<void*> ptrs;
  void* get(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { set_err("cudaMalloc failed"); return nullptr; }
    ptrs.push_back(p);
    return p;
  }
  void* keep(void* p) {   // hand one allocation over to the caller
    ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), p), ptrs.end());
    return p;
  }
  ~ScratchMem() {
    for (void* p : ptrs) cudaFree(p);
  }
};

void destroy_cell_ctx(CellCtx* cc) {
  if (!cc) return;
  if (cc->st) cudaStreamSynchronize(cc->st);
  for (DevBuf* b : {&cc->blobs, &cc->coef, &cc->scalars, &cc->pts, &cc->cells, &cc->proofs, &cc->status, &cc->msm_scratch, &cc->v_commit, &cc->v_cidx, &cc->v_cellidx,
                    &cc->v_cells, &cc->v_proofs, &cc->v_evals, &cc->v_status, &cc->v_pts, &cc->v_wcoef, &cc->v_rpow, &cc->v_scal, &cc->v_r, &cc->v_out,
                    &cc->v_ok, &cc->v_msm_a, &cc->v_msm_b})
    b->release();
  for (void* p : {cc->d_tw, cc->d_naf, cc->d_table, cc->d_prep64, cc->d_fk20})
    if (p) cudaFree(p);
  if (cc->h_stage) cudaFreeHost(cc->h_stage);
  if (cc->ev) cudaEventDestroy(cc->ev);
  for (cudaEvent_t e : cc->ev_v)
    if (e) cudaEventDestroy(e);
  if (cc->st) cudaStreamDestroy(cc->st);
  delete cc;
}

// inverse (natural in -> bit-reversed out, the odd = upper-half outputs dropped) then forward (bit-reversed in ->
// natural out) G1 FFT of size 128 over `batch` vectors: Hhat -> proofs in natural coset order
void cell_g1_idft_dft(CellCtx* cc, void* d_pts, int batch) {
  for (int half = 64; half >= 1; half >>= 1) launch_cell_g1_fft_stage(d_pts, batch, half, true, true, true, cc->d_naf, cc->st);
  for (int half = 1; half <= 64; half <<= 1) launch_cell_g1_fft_stage(d_pts, batch, half, false, false, false, cc->d_naf, cc->st);
}

// Built under c->mu on the first cell call: twiddles, [tau^64]G2 lines, the FK20 points X^b = DFT_128 of the
// reversed strided SRS columns, and their digit table.
bool cell_ctx_build(Ctx* c) {
  if (c->cell) return true;
  if (!c->d_mono) {
    set_err("cell operations need the monomial SRS: load the settings with load_trusted_setup[_file] in this mode");
    return false;
  }
  CellCtx* cc = new CellCtx();
  bool good = [&]() -> bool {
    CU_TRY(cudaStreamCreateWithFlags(&cc->st, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&cc->ev, cudaEventDisableTiming));
    for (cudaEvent_t& e : cc->ev_v) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU_TRY(cudaMalloc(&cc->d_tw, (size_t)EXT_POINTS * 32));
    CU_TRY(cudaMalloc(&cc->d_naf, (size_t)128 * CELL_NAF_BYTES));
    launch_cell_twiddles(cc->d_tw, cc->st);
    launch_cell_twiddle_naf(cc->d_naf, cc->st);
    // g2[64]
    {
      ScratchMem tmp;
      uint32_t g2canon[48];
      const g2_t& q = c->g2_host[CELL_ELEMS];
      blst_fp_to_canon(&g2canon[0], &q.x.fp[0]);
      blst_fp_to_canon(&g2canon[12], &q.x.fp[1]);
      blst_fp_to_canon(&g2canon[24], &q.y.fp[0]);
      blst_fp_to_canon(&g2canon[36], &q.y.fp[1]);
      void* d_g2 = tmp.get(sizeof(g2canon));
      int* d_bad = (int*)tmp.get(sizeof(int));
      if (!d_g2 || !d_bad) return false;
      CU_TRY(cudaMalloc(&cc->d_prep64, g2_prepared_bytes()));
      CU_TRY(cudaMemcpyAsync(d_g2, g2canon, sizeof(g2canon), cudaMemcpyHostToDevice, cc->st));
      launch_g2_prepare(cc->d_prep64, d_bad, d_g2, cc->st);
      int bad = 1;
      CU_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, cc->st));
      CU_TRY(cudaStreamSynchronize(cc->st));
      if (bad) { set_err("g2_values[64] is not on the twist"); return false; }
    }
    // FK20 points
    ScratchMem tmp;
    void* d_pts = tmp.get((size_t)EXT_POINTS * XYZZ_BYTES);
    void* d_aff = tmp.get((size_t)EXT_POINTS * AFFINE_BYTES);
    if (!d_pts || !d_aff) return false;
    launch_cell_srs_columns(d_pts, c->d_mono, cc->st);
    for (int half = 64; half >= 1; half >>= 1) launch_cell_g1_fft_stage(d_pts, 64, half, true, false, false, cc->d_naf, cc->st);
    launch_cell_fk20_points(d_aff, d_pts, cc->st);
    // digit table, window shrunk to what free HBM allows (6 GiB kept for the batch buffers)
    long want;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      want = opts().cell_window_bits;
    }
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    int cbits = (int)std::min(std::max(want, 4L), 14L);
    for (;; cbits--) {
      const size_t need = 2 * cell_table_half_entries(cbits) * AFFINE_BYTES;
      if (need + (size_t(6) << 30) <= free_b || cbits <= 4) break;
    }
    cc->c = cbits;
    cc->nwin = table_num_windows(cbits);
    // one table per half of the frequencies (4096 points each, the commitment table's own layout), back to back
    const size_t half = cell_table_half_entries(cbits);
    CU_TRY(cudaMalloc(&cc->d_table, 2 * half * AFFINE_BYTES));
    void* d_bases = tmp.get((size_t)cc->nwin * N_POINTS * AFFINE_BYTES);
    if (!d_bases) return false;
    for (int v = 0; v < 2; v++) {
      launch_table_bases(d_bases, (const uint8_t*)d_aff + (size_t)v * N_POINTS * AFFINE_BYTES, cc->c, cc->nwin, N_POINTS, cc->st);
      launch_table_fill((uint8_t*)cc->d_table + (size_t)v * half * AFFINE_BYTES, d_bases, cc->c, cc->nwin, N_POINTS, table_top_count(cc->c), cc->st);
    }
    CU_TRY(cudaStreamSynchronize(cc->st));
    CU_TRY(cudaGetLastError());
    cc->d_fk20 = tmp.keep(d_aff);
    return true;
  }();
  if (!good) {
    std::string e = tl_err;
    destroy_cell_ctx(cc);
    set_err(e);
    return false;
  }
  c->cell = cc;
  return true;
}

long cell_chunk() {
  std::lock_guard<std::mutex> lk(g_mu);
  return std::max(1L, opts().cell_chunk_blobs);
}

bool cell_reserve(CellCtx* cc, size_t m, bool own_io) {
  if (own_io && !(cc->blobs.ensure(m * BLOB_BYTES) && cc->cells.ensure(m * N_CELLS * CELL_BYTES) && cc->proofs.ensure(m * N_CELLS * 48))) return false;
  return cc->coef.ensure(m * BLOB_BYTES) && cc->scalars.ensure(m * EXT_POINTS * 32) && cc->pts.ensure(m * N_CELLS * XYZZ_BYTES) &&
         cc->status.ensure(m * sizeof(int)) && cc->msm_scratch.ensure(cell_msm_scratch_bytes((int)m));
}

// coefficient form (cc->coef, m blobs) -> 128 proofs per blob
void cell_enqueue_proofs(CellCtx* cc, void* d_proofs, int m) {
  launch_cell_toeplitz(cc->scalars.p, cc->coef.p, m, cc->d_tw, cc->st);
  launch_cell_msm(cc->pts.p, cc->d_table, cc->c, cc->scalars.p, m, cc->msm_scratch.p, cc->st);
  cell_g1_idft_dft(cc, cc->pts.p, m);
  launch_cell_proofs_finalize(d_proofs, cc->pts.p, m, cc->st);
}

// one chunk, device pointers, on cc->st
bool cell_enqueue_chunk(Ctx* c, const void* d_blobs, int m, void* d_cells, void* d_proofs, int* d_status) {
  CellCtx* cc = c->cell;
  CU_TRY(cudaMemsetAsync(d_status, 0, (size_t)m * sizeof(int), cc->st));
  if (c->lagrange()) launch_le_blob_check(d_status, d_blobs, m, cc->st, c->be_wire());
  launch_cell_poly(cc->coef.p, d_cells, d_blobs, m, c->mode, cc->d_tw, cc->st);
  CU_TRY(cudaEventRecord(cc->ev_v[0], cc->st));   // status and cells are final here; the proofs take 50x longer
  if (d_proofs) cell_enqueue_proofs(cc, d_proofs, m);
  return true;
}

// host buffers, ONE device; status[n] receives one code per blob
C_KZG_RET cells_host_batch_on(Ctx* c, size_t n, const Blob* blobs, Cell* cells_out, KZGProof* proofs_out, int* status) {
  CtxLock lock(c);
  if (!cell_ctx_build(c)) return C_KZG_ERROR;
  CellCtx* cc = c->cell;
  const size_t chunk = std::min<size_t>(n, (size_t)cell_chunk());
  const size_t per = (cells_out ? (size_t)N_CELLS * CELL_BYTES : 0) + (proofs_out ? (size_t)N_CELLS * 48 : 0) + sizeof(int);
  if (!cell_reserve(cc, chunk, true)) return C_KZG_ERROR;
  if (chunk * per > cc->h_stage_cap) {
    if (cc->h_stage) cudaFreeHost(cc->h_stage);
    cc->h_stage = nullptr;
    cc->h_stage_cap = 0;
    if (cudaMallocHost(&cc->h_stage, chunk * per) != cudaSuccess) { set_err("cudaMallocHost failed"); return C_KZG_MALLOC; }
    cc->h_stage_cap = chunk * per;
  }
  C_KZG_RET first = C_KZG_OK;
  for (size_t off = 0; off < n; off += chunk) {
    const size_t m = std::min(chunk, n - off);
    uint8_t* hs = (uint8_t*)cc->h_stage;
    uint8_t* h_cells = hs;
    uint8_t* h_proofs = h_cells + (cells_out ? m * N_CELLS * CELL_BYTES : 0);
    int* h_status = (int*)(h_proofs + (proofs_out ? m * N_CELLS * 48 : 0));
    bool clean = false;   // every item of the pass succeeded: results went straight to the caller's buffers
    bool good = [&]() -> bool {
      CU_TRY(cudaMemcpyAsync(cc->blobs.p, blobs + off, m * BLOB_BYTES, cudaMemcpyHostToDevice, cc->st));
      if (!cell_enqueue_chunk(c, cc->blobs.p, (int)m, cells_out ? cc->cells.p : nullptr, proofs_out ? cc->proofs.p : nullptr, (int*)cc->status.p)) return false;
      // The status array and the cells are complete after the first two kernels of the pass (~2 ms); the proofs
      // follow ~80 ms later.  A side stream fetches the status as soon as it exists -- it decides where the results go:
      // straight into the caller's memory when every item is good (the usual case: no staging copy of 256 KiB per
      // blob on the host), through the pinned staging area when a failed item's slot must stay untouched -- and
      // then copies the cells while the proof kernels run.
      cudaStream_t side = c->slot[2].st;
      CU_TRY(cudaStreamWaitEvent(side, cc->ev_v[0], 0));
      CU_TRY(cudaMemcpyAsync(h_status, cc->status.p, m * sizeof(int), cudaMemcpyDeviceToHost, side));
      CU_TRY(cudaStreamSynchronize(side));
      clean = true;
      for (size_t i = 0; i < m; i++) clean = clean && h_status[i] == 0;
      uint8_t* dst_cells = (clean && cells_out) ? (uint8_t*)(cells_out + off * N_CELLS) : h_cells;
      uint8_t* dst_proofs = (clean && proofs_out) ? (uint8_t*)(proofs_out + off * N_CELLS) : h_proofs;
      if (cells_out) CU_TRY(cudaMemcpyAsync(dst_cells, cc->cells.p, m * N_CELLS * CELL_BYTES, cudaMemcpyDeviceToHost, side));
      if (proofs_out) CU_TRY(cudaMemcpyAsync(dst_proofs, cc->proofs.p, m * N_CELLS * 48, cudaMemcpyDeviceToHost, cc->st));
      CU_TRY(cudaStreamSynchronize(cc->st));
      CU_TRY(cudaStreamSynchronize(side));
      CU_TRY(cudaGetLastError());
      return true;
    }();
    if (!good) return C_KZG_ERROR;
    for (size_t i = 0; i < m; i++) {
      if (h_status[i] == 0) {
        if (!clean) {
          if (cells_out) memcpy(&cells_out[(off + i) * N_CELLS], h_cells + i * N_CELLS * CELL_BYTES, (size_t)N_CELLS * CELL_BYTES);
          if (proofs_out) memcpy(&proofs_out[(off + i) * N_CELLS], h_proofs + i * N_CELLS * 48, (size_t)N_CELLS * 48);
        }
      } else if (first == C_KZG_OK) {
        first = (C_KZG_RET)h_status[i];
      }
      status[off + i] = h_status[i];
    }
  }
  (void)first;
  return C_KZG_OK;
}

// Host-buffer cell batches: one device, or -- after lwkzg_set_devices -- contiguous shards of the blobs on every
// listed GPU (cells and proofs of different blobs are independent: no exchange), one host thread per device.
C_KZG_RET cells_host_batch(const KZGSettings* s, size_t n, const Blob* blobs, Cell* cells_out, KZGProof* proofs_out, int* status) {
  if (n == 0) return C_KZG_OK;
  if (!blobs || (!cells_out && !proofs_out)) { set_err("null argument"); return C_KZG_BADARGS; }
  Ctx* c = ctx_of(s);
  if (!c) return C_KZG_ERROR;
  if (!c->srs_valid) {
    set_err("SRS re-hydration failed: g1_values holds a point that is not on the curve");
    if (status) for (size_t i = 0; i < n; i++) status[i] = C_KZG_ERROR;
    return C_KZG_ERROR;
  }
  std::vector<int> st_own;
  int* st = status;
  if (!st) { st_own.assign(n, 0); st = st_own.data(); }
  C_KZG_RET rc = C_KZG_OK;
  std::vector<Ctx*> ctxs = n >= 2 ? ctxs_for(c) : std::vector<Ctx*>{c};
  if (ctxs.size() > 1 && n >= 2 * ctxs.size()) {
    // replicas are built from the host arrays, which hold the Lagrange points in the Lagrange modes: hand them the
    // monomial points the primary context kept
    for (Ctx* r : ctxs) {
      if (r == c || r->d_mono || !c->d_mono) continue;
      std::lock_guard<std::mutex> lk(r->mu);
      DeviceGuard dg(r->device);
      if (cudaMalloc(&r->d_mono, (size_t)N_POINTS * AFFINE_BYTES) != cudaSuccess ||
          cudaMemcpyPeer(r->d_mono, r->device, c->d_mono, c->device, (size_t)N_POINTS * AFFINE_BYTES) != cudaSuccess) {
        set_err("could not replicate the monomial SRS");
        return C_KZG_ERROR;
      }
      r->mono_owned = true;
    }
    const size_t g = ctxs.size();
    std::vector<C_KZG_RET> rcs(g, C_KZG_OK);
    std::vector<std::string> errs(g);
    std::vector<std::thread> th;
    for (size_t k = 0; k < g; k++)
      th.emplace_back([&, k]() {
        size_t first, cnt;
        shard_of(n, g, k, first, cnt);
        if (!cnt) return;
        rcs[k] = cells_host_batch_on(ctxs[k], cnt, blobs + first, cells_out ? cells_out + first * N_CELLS : nullptr,
                                     proofs_out ? proofs_out + first * N_CELLS : nullptr, st + first);
        if (rcs[k] != C_KZG_OK) errs[k] = tl_err;
      });
    for (auto& t : th) t.join();
    for (size_t k = 0; k < g; k++)
      if (rcs[k] != C_KZG_OK && rc == C_KZG_OK) { rc = rcs[k]; set_err(errs[k]); }
  } else {
    rc = cells_host_batch_on(c, n, blobs, cells_out, proofs_out, st);
  }
  if (rc != C_KZG_OK) return rc;
  if (!status) {
    for (size_t i = 0; i < n; i++)
      if (st[i]) { set_err("invalid blob"); return (C_KZG_RET)st[i]; }
  }
  return C_KZG_OK;
}

C_KZG_RET cells_device_batch(const KZGSettings* s, size_t n, const void* d_blobs, void* d_cells, void* d_proofs, void* d_status, cudaStream_t user) {
  if (n == 0) return C_KZG_OK;
  if (!d_blobs || (!d_cells && !d_proofs)) { set_err("null argument"); return C_KZG_BADARGS; }
  Ctx* c = ctx_of(s);
  if (!c) return C_KZG_ERROR;
  if (!c->srs_valid) { set_err("SRS re-hydration failed"); return C_KZG_ERROR; }
  CtxLock lock(c);
  if (!cell_ctx_build(c)) return C_KZG_ERROR;
  CellCtx* cc = c->cell;
  const size_t chunk = std::min<size_t>(n, (size_t)cell_chunk());
  // growing a buffer frees the old one: wait for whatever still uses it
  if (cc->coef.cap < chunk * BLOB_BYTES || cc->status.cap < chunk * sizeof(int) || cc->msm_scratch.cap < cell_msm_scratch_bytes((int)chunk)) cudaStreamSynchronize(cc->st);
  if (!cell_reserve(cc, chunk, false)) return C_KZG_ERROR;
  bool good = [&]() -> bool {
    CU_TRY(cudaEventRecord(cc->ev, user));
    CU_TRY(cudaStreamWaitEvent(cc->st, cc->ev, 0));
    for (size_t off = 0; off < n; off += chunk) {
      const int m = (int)std::min(chunk, n - off);
      int* st = d_status ? (int*)d_status + off : (int*)cc->status.p;
      void* co = d_cells ? (uint8_t*)d_cells + off * N_CELLS * CELL_BYTES : nullptr;
      void* po = d_proofs ? (uint8_t*)d_proofs + off * N_CELLS * 48 : nullptr;
      if (!cell_enqueue_chunk(c, (const uint8_t*)d_blobs + off * BLOB_BYTES, m, co, po, st)) return false;
      if (c->lagrange()) {
        if (co) launch_zero_failed(co, N_CELLS * CELL_BYTES, st, m, cc->st);
        if (po) launch_zero_failed(po, N_CELLS * 48, st, m, cc->st);
      }
    }
    CU_TRY(cudaEventRecord(cc->ev, cc->st));
    CU_TRY(cudaStreamWaitEvent(user, cc->ev, 0));
    CU_TRY(cudaGetLastError());
    return true;
  }();
  return good ? C_KZG_OK : C_KZG_ERROR;
}

// ------------------------------------------------------------------ verify_cell_kzg_proof_batch
C_KZG_RET cells_verify_batch(bool* ok, const Bytes48* commitments, const uint64_t* cell_indices, const Cell* cells, const Bytes48* proofs, size_t n,
                             const KZGSettings* s) {
  if (!ok) return C_KZG_BADARGS;
  *ok = false;
  Ctx* c = ctx_of(s);
  if (!c) return C_KZG_ERROR;
  const C_KZG_RET bad = c->lagrange() ? C_KZG_BADARGS : C_KZG_ERROR;
  if (n == 0) { *ok = true; return C_KZG_OK; }
  if (!commitments || !cell_indices || !cells || !proofs) { set_err("null argument"); return C_KZG_BADARGS; }
  if (n > (size_t(1) << 24)) { set_err("too many cells"); return C_KZG_BADARGS; }
  if (!c->srs_valid || !c->g2_valid) { set_err("SRS re-hydration failed"); return C_KZG_ERROR; }
  for (size_t i = 0; i < n; i++)
    if (cell_indices[i] >= N_CELLS) { set_err("cell index out of range"); return bad; }
  // deduplicated commitments in order of first appearance (the spec hashes them that way)
  std::vector<Bytes48> uniq;
  std::vector<uint32_t> cidx(n);
  {
    std::map<std::string, uint32_t> seen;
    for (size_t i = 0; i < n; i++) {
      std::string key((const char*)commitments[i].bytes, 48);
      auto it = seen.find(key);
      if (it == seen.end()) {
        it = seen.emplace(key, (uint32_t)uniq.size()).first;
        uniq.push_back(commitments[i]);
      }
      cidx[i] = it->second;
    }
  }
  const size_t nc = uniq.size(), np = n + nc + CELL_ELEMS;
  CtxLock lock(c);
  if (!cell_ctx_build(c)) return C_KZG_ERROR;
  CellCtx* cc = c->cell;
  cudaStream_t st = cc->st;
  std::vector<int> h_status(2 * n + nc);
  int okv = 0;
  bool good = [&]() -> bool {
    if (!(cc->v_commit.ensure(nc * 48) && cc->v_cidx.ensure(n * 4) && cc->v_cellidx.ensure(n * 8) && cc->v_cells.ensure(n * CELL_BYTES) &&
          cc->v_proofs.ensure(n * 48) && cc->v_evals.ensure(n * CELL_BYTES) && cc->v_status.ensure((2 * n + nc) * sizeof(int)) &&
          cc->v_pts.ensure(np * AFFINE_BYTES) && cc->v_wcoef.ensure(n * CELL_BYTES) && cc->v_rpow.ensure(n * 32) && cc->v_scal.ensure(np * 32) &&
          cc->v_r.ensure(32) && cc->v_out.ensure(288) && cc->v_ok.ensure(sizeof(int)) && cc->v_msm_a.ensure(var_msm_scratch_bytes(n)) &&
          cc->v_msm_b.ensure(var_msm_scratch_bytes(np))))
      return false;
    int* d_st = (int*)cc->v_status.p;   // [0, n) cells | [n, 2n) proofs | [2n, 2n + nc) commitments
    uint8_t* pts = (uint8_t*)cc->v_pts.p;
    CU_TRY(cudaMemcpyAsync(cc->v_commit.p, uniq.data(), nc * 48, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(cc->v_cidx.p, cidx.data(), n * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(cc->v_cellidx.p, cell_indices, n * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(cc->v_cells.p, cells, n * CELL_BYTES, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(cc->v_proofs.p, proofs, n * 48, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(d_st, 0, (2 * n + nc) * sizeof(int), st));
    // Three things need nothing but the inputs and run side by side: the challenge r (ONE sequential SHA-256 over
    // 2112 bytes per cell, the longest of them) on the context's high-priority hash stream, the decompression of the
    // proofs on a second stream, cell parsing + commitment decompression here.
    cudaStream_t sh = c->hash_st, sp = c->slot[1].st;
    CU_TRY(cudaEventRecord(cc->ev_v[0], st));
    CU_TRY(cudaStreamWaitEvent(sh, cc->ev_v[0], 0));
    CU_TRY(cudaStreamWaitEvent(sp, cc->ev_v[0], 0));
    launch_cell_batch_challenge(cc->v_r.p, cc->v_commit.p, (int)nc, (const uint32_t*)cc->v_cidx.p, (const uint64_t*)cc->v_cellidx.p, cc->v_cells.p,
                                cc->v_proofs.p, (int)n, c->mode, sh);
    CU_TRY(cudaEventRecord(cc->ev_v[1], sh));
    launch_g1_decompress(pts, nullptr, d_st + n, cc->v_proofs.p, (int)n, sp, c->lagrange());
    CU_TRY(cudaEventRecord(cc->ev_v[2], sp));
    launch_cell_parse(cc->v_evals.p, d_st, cc->v_cells.p, (int)n, c->mode, st);
    launch_g1_decompress(pts + n * AFFINE_BYTES, nullptr, d_st + 2 * n, cc->v_commit.p, (int)nc, st, c->lagrange());
    CU_TRY(cudaMemcpyAsync(pts + (n + nc) * AFFINE_BYTES, c->d_mono, (size_t)CELL_ELEMS * AFFINE_BYTES, cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaStreamWaitEvent(st, cc->ev_v[1], 0));
    launch_cell_verify_scalars(cc->v_wcoef.p, cc->v_rpow.p, cc->v_scal.p, cc->v_evals.p, (const uint64_t*)cc->v_cellidx.p, (const uint32_t*)cc->v_cidx.p,
                               (int)n, (int)nc, cc->v_r.p, cc->d_tw, st);
    CU_TRY(cudaEventRecord(cc->ev_v[3], st));
    // the two MSMs side by side: sum r^k pi_k on the proofs' stream, the right-hand side here
    uint8_t* out = (uint8_t*)cc->v_out.p;
    CU_TRY(cudaStreamWaitEvent(sp, cc->ev_v[3], 0));
    launch_var_msm_mont(out, pts, cc->v_rpow.p, n, cc->v_msm_a.p, sp);
    CU_TRY(cudaEventRecord(cc->ev_v[4], sp));
    CU_TRY(cudaMemsetAsync(out + 96, 0, 96, st));
    CU_TRY(cudaStreamWaitEvent(st, cc->ev_v[2], 0));   // the proofs are decoded
    launch_var_msm_mont(out + 192, pts, cc->v_scal.p, np, cc->v_msm_b.p, st);       // RLP + RLC - RLI
    CU_TRY(cudaStreamWaitEvent(st, cc->ev_v[4], 0));
    launch_batch_final((int*)cc->v_ok.p, out, 1, c->d_prep0, cc->d_prep64, st);
    CU_TRY(cudaMemcpyAsync(h_status.data(), d_st, h_status.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&okv, cc->v_ok.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaGetLastError());
    return true;
  }();
  if (!good) return C_KZG_ERROR;
  for (int v : h_status)
    if (v) { set_err("invalid cell, proof or commitment"); return c->lagrange() ? bad_code(v) : C_KZG_ERROR; }
  *ok = okv != 0;
  return C_KZG_OK;
}

// ------------------------------------------------------------------ recover_cells_and_kzg_proofs
C_KZG_RET cells_recover(Cell* cells_out, KZGProof* proofs_out, const uint64_t* cell_indices, const Cell* cells, size_t n, const KZGSettings* s) {
  if (!cell_indices || !cells || (!cells_out && !proofs_out)) { set_err("null argument"); return C_KZG_BADARGS; }
  Ctx* c = ctx_of(s);
  if (!c) return C_KZG_ERROR;
  const C_KZG_RET bad = c->lagrange() ? C_KZG_BADARGS : C_KZG_ERROR;
  if (n < N_CELLS / 2 || n > N_CELLS) { set_err("need between 64 and 128 cells"); return bad; }
  for (size_t i = 0; i < n; i++) {
    if (cell_indices[i] >= N_CELLS) { set_err("cell index out of range"); return bad; }
    if (i && cell_indices[i] <= cell_indices[i - 1]) { set_err("cell indices must be strictly ascending"); return bad; }
  }
  if (!c->srs_valid) { set_err("SRS re-hydration failed"); return C_KZG_ERROR; }
  CtxLock lock(c);
  if (!cell_ctx_build(c)) return C_KZG_ERROR;
  CellCtx* cc = c->cell;
  cudaStream_t st = cc->st;
  std::vector<int> h_status(n + 1);
  bool good = [&]() -> bool {
    if (!cell_reserve(cc, 1, true)) return false;
    // coefficient output + 3 x 8192 field elements of workspace
    if (!(cc->coef.ensure((size_t)(4096 + 3 * 8192) * 32) && cc->v_cellidx.ensure(n * 8) && cc->v_cells.ensure(n * CELL_BYTES) &&
          cc->v_evals.ensure(n * CELL_BYTES) && cc->v_status.ensure((n + 1) * sizeof(int))))
      return false;
    int* d_st = (int*)cc->v_status.p;
    CU_TRY(cudaMemcpyAsync(cc->v_cellidx.p, cell_indices, n * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(cc->v_cells.p, cells, n * CELL_BYTES, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(d_st, 0, (n + 1) * sizeof(int), st));
    launch_cell_parse(cc->v_evals.p, d_st, cc->v_cells.p, (int)n, c->mode, st);
    launch_cell_recover(cc->coef.p, d_st + n, cc->v_evals.p, (const uint64_t*)cc->v_cellidx.p, (int)n, cc->d_tw, st);
    if (cells_out) launch_cell_coef_to_cells(cc->cells.p, cc->coef.p, 1, c->mode, cc->d_tw, st);
    if (proofs_out) cell_enqueue_proofs(cc, cc->proofs.p, 1);
    CU_TRY(cudaMemcpyAsync(h_status.data(), d_st, h_status.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaGetLastError());
    return true;
  }();
  if (!good) return C_KZG_ERROR;
  for (int v : h_status)
    if (v) { set_err("invalid or inconsistent cells"); return c->lagrange() ? bad_code(v) : C_KZG_ERROR; }
  // outputs are copied only now that every check has passed
  if (cells_out && cudaMemcpy(cells_out, cc->cells.p, (size_t)N_CELLS * CELL_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) { set_err("D2H failed"); return C_KZG_ERROR; }
  if (proofs_out && cudaMemcpy(proofs_out, cc->proofs.p, (size_t)N_CELLS * 48, cudaMemcpyDeviceToHost) != cudaSuccess) { set_err("D2H failed"); return C_KZG_ERROR; }
  return C_KZG_OK;
}

// Test hook: the intermediate values of the FK20 pipeline for ONE blob, so that a parity test can say WHICH stage
// deviates: the 8192 MSM scalars (canonical little-endian limbs, index (j / 64) * 4096 + b * 64 + j % 64), Hhat_j (compressed, natural j), H after the
// inverse FFT (compressed, position p holds H_brp7(p), odd positions infinity) and the FK20 points in the same order
// (canonical little-endian limbs x || y, 96 bytes each).
C_KZG_RET cells_debug_stages(uint8_t* scalars, uint8_t* hhat48, uint8_t* h48, uint8_t* fk20_xy96, const Blob* blob, const KZGSettings* s) {
  Ctx* c = ctx_of(s);
  if (!c || !c->srs_valid) return C_KZG_ERROR;
  CtxLock lock(c);
  if (!cell_ctx_build(c)) return C_KZG_ERROR;
  CellCtx* cc = c->cell;
  cudaStream_t st = cc->st;
  bool good = [&]() -> bool {
    if (!cell_reserve(cc, 1, true)) return false;
    struct Tmp : DevBuf { ~Tmp() { release(); } } tmp;
    if (!tmp.ensure((size_t)EXT_POINTS * 48)) return false;
    CU_TRY(cudaMemcpyAsync(cc->blobs.p, blob, BLOB_BYTES, cudaMemcpyHostToDevice, st));
    launch_cell_poly(cc->coef.p, nullptr, cc->blobs.p, 1, c->mode, cc->d_tw, st);
    launch_cell_toeplitz(cc->scalars.p, cc->coef.p, 1, cc->d_tw, st);
    CU_TRY(cudaMemcpyAsync(scalars, cc->scalars.p, (size_t)EXT_POINTS * 32, cudaMemcpyDeviceToHost, st));
    launch_cell_msm(cc->pts.p, cc->d_table, cc->c, cc->scalars.p, 1, cc->msm_scratch.p, st);
    launch_msm_finalize(tmp.p, nullptr, cc->pts.p, 1, N_CELLS, st);
    CU_TRY(cudaMemcpyAsync(hhat48, tmp.p, (size_t)N_CELLS * 48, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (int half = 64; half >= 1; half >>= 1) launch_cell_g1_fft_stage(cc->pts.p, 1, half, true, true, true, cc->d_naf, st);
    launch_msm_finalize(tmp.p, nullptr, cc->pts.p, 1, N_CELLS, st);
    CU_TRY(cudaMemcpyAsync(h48, tmp.p, (size_t)N_CELLS * 48, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    // the FK20 points as canonical little-endian limbs x || y (96 bytes each)
    if (!tmp.ensure((size_t)EXT_POINTS * 96)) return false;
    launch_affine_to_canon(tmp.p, cc->d_fk20, EXT_POINTS, st);
    CU_TRY(cudaMemcpyAsync(fk20_xy96, tmp.p, (size_t)EXT_POINTS * 96, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaGetLastError());
    return true;
  }();
  return good ? C_KZG_OK : C_KZG_ERROR;
}
