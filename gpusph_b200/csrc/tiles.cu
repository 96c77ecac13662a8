// tiles.cu — work decomposition for the shared-memory-staged forces kernel.
//
// No counterpart in the reference (its forces kernels gather every neighbour through the texture/L1 path).
// Because COORD1 is the fastest digit of the cell hash (src/cuda/cellgrid.cuh:101-106), the particles of a run
// of cells along COORD1 inside one (COORD2, COORD3) row are ONE contiguous range of the sorted arrays, and the
// 27-cell neighbourhood of such a run is 9 contiguous ranges (one per neighbouring row). A tile is a run of
// consecutive non-empty cells with at most g.tile_p particles whose 9 ranges hold at most g.tile_s particles;
// the forces kernel stages those ranges into shared memory with bulk (TMA) copies.
#include "common.cuh"
#include <cub/device/device_scan.cuh>

struct RowGeom {
	int G1, G2, G3;        // grid size along COORD1,2,3
	int per2, per3;        // periodicity along COORD2, COORD3
	uint tile_p, tile_s;   // max central particles per pass / max staged particles of the forces kernel configuration
};

__device__ __forceinline__ uint cell_count(const uint *__restrict__ cs, const uint *__restrict__ ce, int cell)
{
	const uint s = __ldg(cs + cell);
	return s == CELL_EMPTY ? 0u : __ldg(ce + cell) - s;
}

// neighbouring row (c2+d2, c3+d3) -> first cell hash of that row, or -1 if outside a non-periodic domain
__device__ __forceinline__ int neib_row_base(const RowGeom &g, int c2, int c3, int d2, int d3)
{
	int n2 = c2 + d2, n3 = c3 + d3;
	if (n2 < 0) { if (g.per2) n2 = g.G2 - 1; else return -1; }
	if (n2 >= g.G2) { if (g.per2) n2 = 0; else return -1; }
	if (n3 < 0) { if (g.per3) n3 = g.G3 - 1; else return -1; }
	if (n3 >= g.G3) { if (g.per3) n3 = 0; else return -1; }
	return (n3 * g.G2 + n2) * g.G1;
}

// One thread per row of cells. WRITE=false: count the tiles of the row. WRITE=true: emit them at row_offset[row].
template<bool WRITE>
__global__ void __launch_bounds__(128)
build_tiles_kernel(const RowGeom g, const uint *__restrict__ cellStart, const uint *__restrict__ cellEnd,
	uint *__restrict__ row_tiles, Tile *__restrict__ tiles, uint *__restrict__ tile_info, const uint range_end)
{
	const int row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= g.G2 * g.G3) return;
	const int c2 = row % g.G2, c3 = row / g.G2;
	int rbase[9];
#pragma unroll
	for (int r = 0; r < 9; ++r) rbase[r] = neib_row_base(g, c2, c3, r % 3 - 1, r / 3 - 1);
	const int base = row * g.G1;

	auto col = [&](int c) -> uint {      // particles in column c of the 9 rows
		if (c < 0 || c >= g.G1) return 0u;
		uint t = 0;
#pragma unroll
		for (int r = 0; r < 9; ++r) if (rbase[r] >= 0) t += cell_count(cellStart, cellEnd, rbase[r] + c);
		return t;
	};
	auto emit = [&](int a, int b, uint slot) {
		Tile t;
		t.first = __ldg(cellStart + base + a);
		uint end = __ldg(cellEnd + base + b);
		if (end > range_end) end = range_end;
		t.count = end - t.first;
		const int lo = max(a - 1, 0), hi = min(b + 1, g.G1 - 1);
		uint staged = 0;
#pragma unroll
		for (int r = 0; r < 9; ++r) {
			uint s = 0, e = 0;
			if (rbase[r] >= 0) {
				for (int c = lo; c <= hi; ++c) {
					const uint cs = __ldg(cellStart + rbase[r] + c);
					if (cs != CELL_EMPTY) { if (e == 0) s = cs; e = __ldg(cellEnd + rbase[r] + c); }
				}
			}
			t.row_start[r] = s;
			t.row_count[r] = e ? e - s : 0u;
			staged += t.row_count[r];
		}
		if (staged > g.tile_s) tile_info[1] = 1;      // cannot be staged: the caller falls back to the gather kernel
		tiles[slot] = t;
	};

	uint ntiles = 0;
	const uint out = WRITE ? row_tiles[row] : 0u;
	bool open = false;
	int a = 0, b = 0;
	uint central = 0, staged_inner = 0;
	uint col_prev = 0, col_cur = col(0), col_next;
	for (int c = 0; c < g.G1; ++c) {
		col_next = col(c + 1);
		uint cnt = 0;
		const uint cs = __ldg(cellStart + base + c);
		if (cs != CELL_EMPTY && cs < range_end) cnt = __ldg(cellEnd + base + c) - cs;
		const bool full = open && (cnt == 0 || central + cnt > g.tile_p || staged_inner + col_cur + col_next > g.tile_s);
		if (full) {
			if (WRITE) emit(a, b, out + ntiles);
			++ntiles;
			open = false;
		}
		if (cnt) {
			if (!open) { open = true; a = c; central = 0; staged_inner = col_prev; }
			central += cnt; staged_inner += col_cur; b = c;
		}
		col_prev = col_cur; col_cur = col_next;
	}
	if (open) { if (WRITE) emit(a, b, out + ntiles); ++ntiles; }
	if (!WRITE) row_tiles[row] = ntiles;
}

__global__ void finish_tile_scan_kernel(const uint *__restrict__ row_tiles_excl, const uint *__restrict__ row_counts_last,
	uint *__restrict__ tile_info, const int rows)
{
	// total = exclusive[rows-1] + count[rows-1]; counts were overwritten by the scan, so the caller passes the last count
	tile_info[0] = row_tiles_excl[rows - 1] + row_counts_last[0];
}

void b200_invalidate_tiles(b200sph_ctx *ctx) { ctx->tiles_state = 0; ctx->num_tiles = 0; }

int b200_build_tiles(b200sph_ctx *ctx, const uint32_t *cell_start, const uint32_t *cell_end, uint range_end)
{
	b200_invalidate_tiles(ctx);
	const b200sph_params &hp = ctx->hp;
	if (!ctx->use_tiles) return B200SPH_OK;
	if (hp.periodic & (1u << hp.coord[0])) return B200SPH_OK;      // runs along a periodic COORD1 are not contiguous
	RowGeom g;
	g.G1 = (int)hp.grid_size[hp.coord[0]]; g.G2 = (int)hp.grid_size[hp.coord[1]]; g.G3 = (int)hp.grid_size[hp.coord[2]];
	g.per2 = (hp.periodic >> hp.coord[1]) & 1; g.per3 = (hp.periodic >> hp.coord[2]) & 1;
	g.tile_p = (uint)ctx->tile_p; g.tile_s = (uint)ctx->tile_s;
	const int rows = g.G2 * g.G3;
	cudaStream_t s = ctx->stream;
	if (ctx->row_tiles_cap < (size_t)rows + 2) {
		cudaFree(ctx->row_tiles); ctx->row_tiles = NULL; ctx->row_tiles_cap = 0;
		CUDA_TRY(cudaMalloc(&ctx->row_tiles, ((size_t)rows + 2) * 2 * sizeof(uint)));
		ctx->row_tiles_cap = (size_t)rows + 2;
	}
	// a tile holds at least one non-empty cell, hence at least one particle below range_end
	const size_t ncells = (size_t)g.G1 * rows;
	const size_t max_tiles = (ncells < range_end ? ncells : range_end) + 1;
	if (ctx->tiles_cap < max_tiles) {
		cudaFree(ctx->tiles); ctx->tiles = NULL; ctx->tiles_cap = 0;
		CUDA_TRY(cudaMalloc(&ctx->tiles, max_tiles * sizeof(Tile)));
		ctx->tiles_cap = max_tiles;
	}
	uint *counts = ctx->row_tiles, *last = ctx->row_tiles + ctx->row_tiles_cap;
	CUDA_TRY(cudaMemsetAsync(ctx->d_tile_info, 0, 4 * sizeof(uint), s));
	const int nb = (rows + 127) / 128;
	build_tiles_kernel<false><<<nb, 128, 0, s>>>(g, cell_start, cell_end, counts, NULL, ctx->d_tile_info, range_end);
	KERNEL_TRY();
	CUDA_TRY(cudaMemcpyAsync(last, counts + rows - 1, sizeof(uint), cudaMemcpyDeviceToDevice, s));
	size_t tmp = 0;
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(NULL, tmp, counts, counts, rows, s));
	if (ctx->sort_tmp_bytes < tmp) {
		cudaFree(ctx->sort_tmp); ctx->sort_tmp = NULL; ctx->sort_tmp_bytes = 0;
		CUDA_TRY(cudaMalloc(&ctx->sort_tmp, tmp));
		ctx->sort_tmp_bytes = tmp;
	}
	tmp = ctx->sort_tmp_bytes;
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(ctx->sort_tmp, tmp, counts, counts, rows, s));
	finish_tile_scan_kernel<<<1, 1, 0, s>>>(counts, last, ctx->d_tile_info, rows);
	KERNEL_TRY();
	build_tiles_kernel<true><<<nb, 128, 0, s>>>(g, cell_start, cell_end, counts, ctx->tiles, ctx->d_tile_info, range_end);
	KERNEL_TRY();
	CUDA_TRY(cudaMemcpyAsync(ctx->h_tile_info, ctx->d_tile_info, 2 * sizeof(uint), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaEventRecord(ctx->tiles_event, s));
	ctx->tiles_state = 1;
	ctx->tiles_range_end = range_end;
	ctx->tiles_cellstart = cell_start;
	return B200SPH_OK;
}
