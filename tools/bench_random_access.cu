// bench_random_access.cu -- what the memory system of this GPU gives to dependent random 32-byte-sector loads, the access
// pattern of an FM-index walk (K1): one chain's latency and the aggregate rate of many chains, versus footprint.
// The streaming HBM peak (MEASURED_PEAKS.json) is not reachable by this pattern; this is the ceiling K1 is up against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bra tools/bench_random_access.cu && /tmp/bra
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void chase(const uint32_t *a, int steps, uint32_t *out, long long *cyc, uint32_t mask)
{
	uint32_t p = 12345u & mask; long long t0 = clock64();
	for (int i = 0; i < steps; i++) p = mix(p + __ldg(a + p) + i) & mask;
	long long t1 = clock64(); out[0] = p; cyc[0] = t1 - t0;
}
__global__ void chase_many(const uint32_t *a, int steps, uint32_t *out, uint32_t mask)
{
	uint32_t p = mix(blockIdx.x * blockDim.x + threadIdx.x) & mask;
	for (int i = 0; i < steps; i++) p = mix(p + __ldg(a + p) + i) & mask;
	if (p == 0xffffffff) out[0] = p;
}
int main()
{
	uint32_t *out; long long *cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	printf("{\"sms\": %d, \"rows\": [\n", sms);
	for (int lg = 24; lg <= 33; lg++) { // footprint 2^lg bytes
		uint64_t n = (1ull << lg) / 4; uint32_t mask = (uint32_t)(n - 1);
		uint32_t *a; if (cudaMalloc(&a, n * 4) != cudaSuccess) break;
		cudaMemset(a, 0, n * 4); cudaDeviceSynchronize();
		chase<<<1, 1>>>(a, 4000, out, cyc, mask); cudaDeviceSynchronize();
		long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
		printf(" {\"footprint_mb\": %.0f, \"one_chain_cycles_per_load\": %.0f", (double)(1ull << lg) / 1e6, (double)c / 4000);
		for (int wps : {8, 16, 32, 64}) {
			cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
			int steps = 1000, blocks = sms * wps / 4, threads = 128;
			chase_many<<<blocks, threads>>>(a, 100, out, mask);
			cudaEventRecord(e0); chase_many<<<blocks, threads>>>(a, steps, out, mask); cudaEventRecord(e1); cudaDeviceSynchronize();
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			printf(", \"gloads_per_s_%dwarps_per_sm\": %.1f", wps, (double)blocks * threads * steps / (ms * 1e-3) / 1e9);
		}
		printf("}%s\n", lg < 33 ? "," : "");
		cudaFree(a);
	}
	printf("]}\n");
	return 0;
}
