// debug harness: run fast_blur_kernel on a raw u8 image, dump per-pixel strengths around a point
#include "../../srrg2_proslam_b200/csrc/k_detect.cu"
#include <vector>
__global__ void dbg_strength(const uint8_t* img, int stride, int x, int y, int* out) {
  __shared__ uint8_t s[40*72];
  for (int i = threadIdx.x; i < 40*72; i += blockDim.x) { int ty = i/72, tx = i%72; s[i] = img[(y-20+ty)*stride + (x-36+tx)]; }
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = fast_strength(s + 20*72 + 36); }
}
int main(int argc, char** argv) {
  int rows = atoi(argv[2]), cols = atoi(argv[3]); int px = atoi(argv[4]), py = atoi(argv[5]);
  std::vector<uint8_t> h(rows*cols); FILE* f = fopen(argv[1], "rb"); fread(h.data(),1,h.size(),f); fclose(f);
  uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d,h.data(),h.size(),cudaMemcpyHostToDevice);
  int* o; cudaMalloc(&o, 4); dbg_strength<<<1,128>>>(d, cols, px, py, o); int s; cudaMemcpy(&s,o,4,cudaMemcpyDeviceToHost);
  printf("device fast_strength(%d,%d) = %d  err=%s\n", px, py, s, cudaGetErrorString(cudaGetLastError()));
  // full kernel
  int pitch = (cols+127)/128*128; uint8_t *nm,*bl; cudaMalloc(&nm,(size_t)pitch*rows); cudaMalloc(&bl,(size_t)pitch*rows);
  cudaMemset(nm,0,(size_t)pitch*rows);
  dim3 grid((cols+TW-1)/TW,(rows+TH-1)/TH,1);
  fast_blur_kernel<<<grid,K1_THREADS>>>(d,0,rows,cols,cols,15,0,nm,bl,pitch,(long long)pitch*rows);
  std::vector<uint8_t> hn((size_t)pitch*rows); cudaMemcpy(hn.data(),nm,hn.size(),cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  long cnt=0; for(int y=0;y<rows;y++)for(int x=0;x<cols;x++) cnt+= hn[(size_t)y*pitch+x]!=0; printf("corners(thr15,nms0)=%ld\n",cnt);
  for(int y=py-2;y<=py+2;y++){for(int x=px-2;x<=px+2;x++)printf("%4d",hn[(size_t)y*pitch+x]);printf("\n");}
  return 0;
}
