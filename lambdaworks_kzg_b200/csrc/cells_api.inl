// C ABI of the PeerDAS / EIP-7594 cell calls (included by lwkzg.cu inside its extern "C" block; include/lwkzg.h part 3).
C_KZG_RET compute_cells_and_kzg_proofs(Cell* cells, KZGProof* proofs, const Blob* blob, const KZGSettings* s) {
  return cells_host_batch(s, 1, blob, cells, proofs, nullptr);
}
C_KZG_RET lwkzg_compute_cells_and_kzg_proofs_batch(Cell* cells, KZGProof* proofs, const Blob* blobs, size_t n, const KZGSettings* s, int* status) {
  return cells_host_batch(s, n, blobs, cells, proofs, status);
}
C_KZG_RET lwkzg_compute_cells_and_kzg_proofs_batch_device(void* d_cells, void* d_proofs, const void* d_blobs, size_t n, const KZGSettings* s, void* stream,
                                                          void* d_status) {
  return cells_device_batch(s, n, d_blobs, d_cells, d_proofs, d_status, (cudaStream_t)stream);
}
C_KZG_RET recover_cells_and_kzg_proofs(Cell* recovered_cells, KZGProof* recovered_proofs, const uint64_t* cell_indices, const Cell* cells, size_t num_cells,
                                       const KZGSettings* s) {
  return cells_recover(recovered_cells, recovered_proofs, cell_indices, cells, num_cells, s);
}
C_KZG_RET verify_cell_kzg_proof_batch(bool* ok, const Bytes48* commitments_bytes, const uint64_t* cell_indices, const Cell* cells, const Bytes48* proofs_bytes,
                                      size_t num_cells, const KZGSettings* s) {
  return cells_verify_batch(ok, commitments_bytes, cell_indices, cells, proofs_bytes, num_cells, s);
}
int lwkzg_cell_window_bits(const KZGSettings* s) {
  Ctx* c = ctx_of(s);
  if (!c) return -1;
  std::lock_guard<std::mutex> lk(c->mu);
  return c->cell ? c->cell->c : -1;
}
C_KZG_RET lwkzg_debug_cell_stages(uint8_t* scalars, uint8_t* hhat48, uint8_t* h48, uint8_t* fk20_xy96, const Blob* blob, const KZGSettings* s) {
  return cells_debug_stages(scalars, hhat48, h48, fk20_xy96, blob, s);
}
