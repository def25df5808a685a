// kg_snps.cuh -- the SNP twin of the association scan.
//
// Reference: MultipleSNPsDataBases (/root/reference/src/snps_multiple_databases.cpp): the PLINK .bed genotypes of every
// SNP become three bit planes over the phenotyped samples -- presence (A/A), non-missing, heterozygous (ctor :63-135) --
// and get_most_associated_snps (:225-236) scores every SNP with calculate_grammmar_approx_association (:143-158):
//   three dot products of the permuted phenotype vector with the planes in the SSE4 lane order of the k-mer score
//   (dot_product_SSE4 :41-61: 4 float lanes, float sum ((l0+l1)+l2)+l3), then
//   yigi = dot(pa) + 0.5 dot(het);  ss = dot(nonmissing);  N = #non-missing;  S = sum g;  S2 = sum g^2 (g in {0, 1/2, 1})
//   score = (N yigi - S ss)^2 / (N (N S2 - S S)), 0 when mac > S or mac > N - S.
// Every operation of the double epilogue is individually rounded (the reference build has no FMA).
#pragma once
#include "kg_common.cuh"
#include "kg_scan_exact.cuh"

// .bed rows (2 bits per sample of the file, 4 samples per byte) -> planes [n_snps][w_mem] x 3 and the per-SNP sums
__global__ void kg_snp_planes_kernel(const uint8_t *__restrict__ bed, uint64_t n_snps, uint32_t bytes_per_snp,
                                     const uint32_t *__restrict__ map_byte, const uint32_t *__restrict__ map_shift, uint32_t n_samples,
                                     uint32_t w_mem, uint64_t *__restrict__ pa, uint64_t *__restrict__ nonmissing, uint64_t *__restrict__ het,
                                     double *__restrict__ s_g, double *__restrict__ s_n, double *__restrict__ s_g2) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_snps; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint8_t *row = bed + i * bytes_per_snp;
		uint32_t n_pa = 0, n_het = 0, n_tot = 0;
		for (uint32_t w = 0; w < w_mem; w++) {
			uint64_t vp = 0, vn = 0, vh = 0;
			for (uint32_t b = 0; b < 64; b++) {
				const uint32_t si = w * 64 + b;
				if (si >= n_samples) break;
				const uint32_t dubit = (row[map_byte[si]] >> map_shift[si]) & 3u;   // 00 a/a, 01 missing, 10 A/a, 11 A/A
				vp |= (uint64_t)(dubit == 3u) << b;
				vn |= (uint64_t)(dubit != 1u) << b;
				vh |= (uint64_t)(dubit == 2u) << b;
			}
			pa[i * w_mem + w] = vp;
			nonmissing[i * w_mem + w] = vn;
			het[i * w_mem + w] = vh;
			n_pa += __popcll(vp);
			n_het += __popcll(vh);
			n_tot += __popcll(vn);
		}
		// the reference adds 1 / 0.5 per sample in double: exact, so the order does not matter
		s_g[i] = (double)n_pa + 0.5 * (double)n_het;
		s_g2[i] = (double)n_pa + 0.25 * (double)n_het;
		s_n[i] = (double)n_tot;
	}
}

// four adjacent lanes (the SSE lanes) per SNP, one phenotype per pass over y (staged in shared memory in lane order)
__global__ void __launch_bounds__(256) kg_snp_scores_kernel(const uint64_t *__restrict__ pa, const uint64_t *__restrict__ nonmissing,
                                                            const uint64_t *__restrict__ het, const double *__restrict__ s_g,
                                                            const double *__restrict__ s_n, const double *__restrict__ s_g2, uint64_t n_snps,
                                                            uint32_t nb, const float *__restrict__ y_lane /*[P][nb * 128]*/, uint32_t n_pheno,
                                                            double mac, double *__restrict__ scores /*[P][n_snps]*/) {
	extern __shared__ float ys[];   // [nb * 128] lane order: ys[(4 b + L) * 32 + t]
	const uint32_t lane = threadIdx.x & 31, L = lane & 3, grp_base = lane & ~3u;
	const uint32_t w_mem = 2 * nb;
	for (uint32_t p = 0; p < n_pheno; p++) {
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < nb * 128; i += blockDim.x) ys[i] = y_lane[(size_t)p * nb * 128 + i];
		__syncthreads();
		const uint64_t per_pass = ((uint64_t)gridDim.x * blockDim.x) >> 2;
		for (uint64_t i0 = 0; i0 < n_snps; i0 += per_pass) {
			const uint64_t i = i0 + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2);
			const bool valid = i < n_snps;
			float a_pa = 0.f, a_nm = 0.f, a_het = 0.f;
			for (uint32_t b = 0; b < nb; b++) {
				const uint32_t g = 4 * b + L;   // 32-bit half g of the row: word g / 2, half g % 2
				uint32_t wp = 0, wn = 0, wh = 0;
				if (valid) {
					wp = reinterpret_cast<const uint32_t *>(pa + i * w_mem)[g];
					wn = reinterpret_cast<const uint32_t *>(nonmissing + i * w_mem)[g];
					wh = reinterpret_cast<const uint32_t *>(het + i * w_mem)[g];
				}
				const float *yg = ys + g * 32;
#pragma unroll 8
				for (int t = 0; t < 32; t++) {
					const float yv = yg[t];
					const uint32_t m = 0x80000000u >> t;
					float t1[1] = {a_pa}, t2[1] = {a_nm}, t3[1] = {a_het};
					const float yy[1] = {yv};
					kg_pred_add<1>(t1, yy, wp & m);
					kg_pred_add<1>(t2, yy, wn & m);
					kg_pred_add<1>(t3, yy, wh & m);
					a_pa = t1[0]; a_nm = t2[0]; a_het = t3[0];
				}
			}
			float d[3];
			const float acc[3] = {a_pa, a_nm, a_het};
#pragma unroll
			for (int k = 0; k < 3; k++) {
				const float l0 = __shfl_sync(0xffffffffu, acc[k], grp_base + 0);
				const float l1 = __shfl_sync(0xffffffffu, acc[k], grp_base + 1);
				const float l2 = __shfl_sync(0xffffffffu, acc[k], grp_base + 2);
				const float l3 = __shfl_sync(0xffffffffu, acc[k], grp_base + 3);
				d[k] = __fadd_rn(__fadd_rn(__fadd_rn(l0, l1), l2), l3);   // (sumsf[0] + sumsf[1] + sumsf[2] + sumsf[3]) in float
			}
			if (valid && L == 0) {
				const double N = s_n[i], S = s_g[i], S2 = s_g2[i];
				double score = 0.0;
				if (!((mac > S) || (mac > __dsub_rn(N, S)))) {
					const double yigi = __dadd_rn((double)d[0], __dmul_rn((double)d[2], 0.5));
					const double ss = (double)d[1];
					double r = __dsub_rn(__dmul_rn(N, yigi), __dmul_rn(S, ss));
					r = __dmul_rn(r, r);
					const double den = __dmul_rn(N, __dsub_rn(__dmul_rn(N, S2), __dmul_rn(S, S)));
					score = __ddiv_rn(r, den);
				}
				scores[(size_t)p * n_snps + i] = score;
			}
		}
	}
}
