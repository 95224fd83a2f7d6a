#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <string.h>
#include "../mbf_bam_b200/csrc/inflate_core.cuh"
using namespace mbf;
long nlit=0,nmatch=0,matchbytes=0, litlen_hist[16]={0}, dist_near=0, dist_mid=0, dist_far=0, nblocks=0, nsym=0, len_hist[260]={0};
struct CountOut { uint8_t* base; uint32_t pos, cap;
 bool want_flush() const { return pos+258>cap; }
 void put(uint8_t b){ base[pos++]=b; nlit++; }
 bool copy(uint32_t d, uint32_t l){ if(d>pos) return false; for(uint32_t i=0;i<l;i++) base[pos+i]=base[pos-d+i]; pos+=l; nmatch++; matchbytes+=l; len_hist[l]++; if(d<=512) dist_near++; else if (d<=2048) dist_mid++; else dist_far++; return true; } };
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); std::vector<uint8_t> d(n+16); fread(d.data(),1,n,f);
  long off=0; std::vector<uint8_t> out(70000); DecoderTables T; long maxblocks=atol(argv[2]); long skip=atol(argv[3]);
  long bits_lit=0;
  while(off+18<=n && nblocks<maxblocks+skip){
    uint32_t xlen=d[off+10]|(d[off+11]<<8); uint32_t csize=(d[off+16]|(d[off+17]<<8))+1u; uint32_t isize; memcpy(&isize,&d[off+csize-4],4);
    const uint8_t*p=&d[off+12+xlen]; uint32_t clen=csize-12-xlen-8;
    if (nblocks>=skip) {
    BitReader br; br.init(p,clen); CountOut o{out.data(),0,isize+258};
    for(;;){ br.refill(); uint32_t bf=br.take(1),bt=br.take(2);
      if(bt==0){ br.consume(br.cnt&7); br.refill(); uint32_t len=br.take(16); br.refill(); br.take(16); for(uint32_t i=0;i<len;i++){br.refill(); o.base[o.pos++]=br.take(8);} }
      else { if(bt==1) build_fixed(&T); else build_dynamic(&T,br);
        // histogram of code lengths weighted by use: decode manually
        for(;;){ br.refill(); uint32_t e=T.ll_lut[br.peek(LL_ROOT)]; if(e>=LUT_SLOW) e=slow_decode(&T.ll_hd,T.ll_sym,br.peek(15),LL_ROOT); br.consume(e&15); uint32_t sym=e>>4; nsym++;
          if(sym<256){ o.put(sym); litlen_hist[e&15]++; bits_lit+=e&15; continue;} if(sym==256) break; sym-=257; uint32_t len=H_LEN_BASE[sym]+br.take(H_LEN_EXTRA[sym]); br.refill(); e=T.d_lut[br.peek(D_ROOT)]; if(e>=LUT_SLOW) e=slow_decode(&T.d_hd,T.d_sym,br.peek(15),D_ROOT); br.consume(e&15); uint32_t ds=e>>4; uint32_t dist=H_DIST_BASE[ds]+br.take(H_DIST_EXTRA[ds]); o.copy(dist,len);} }
      if(bf) break; }
    }
    off+=csize; nblocks++; }
  printf("blocks %ld lits %ld matches %ld matchbytes %ld avg_len %.2f  lit_bits_avg %.2f\n", nblocks-skip,nlit,nmatch,matchbytes,(double)matchbytes/nmatch,(double)bits_lit/nlit);
  printf("literal code length hist: "); for(int i=1;i<16;i++) printf("%d:%.3f ",i,(double)litlen_hist[i]/nlit); printf("\n");
  printf("dist <=512 %.3f <=2048 %.3f far %.3f\n",(double)dist_near/nmatch,(double)dist_mid/nmatch,(double)dist_far/nmatch);
  printf("match len hist: "); for(int i=3;i<=20;i++) printf("%d:%.3f ",i,(double)len_hist[i]/nmatch); printf("\n");
  printf("symbols per output byte %.3f; out bytes %ld\n",(double)(nlit+nmatch)/(nlit+matchbytes), nlit+matchbytes);
}
