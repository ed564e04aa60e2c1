// cuda_emu.cpp -- TEST INFRASTRUCTURE (see cuda_emu.h): the fibre scheduler behind the CUDA emulator.
#include "cuda_emu.h"
#undef blockIdx
#undef blockDim
#undef gridDim
#undef threadIdx
#include <sys/mman.h>
#include <signal.h>
#include <execinfo.h>
#include <unistd.h>

namespace rb2emu {

thread_local Cta *cta = 0;
thread_local dim3 cur_threadIdx;

// ---- context switch (x86-64 System V): callee-saved registers + stack pointer ---------------------
extern "C" void rb2emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl rb2emu_switch
.type rb2emu_switch,@function
rb2emu_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size rb2emu_switch,.-rb2emu_switch
)");

static const size_t STACK = 256 << 10;

static void fibre_main()
{
	Cta *c = cta;
	Fibre *f = c->cur;
	(*c->body)();
	f->done = true;
	--c->live;
	++c->progress;
	rb2emu_switch(&f->sp, c->sched_sp);
	abort(); // never resumed
}

static void fibre_init(Fibre &f)
{
	// initial frame: six zeroed callee-saved registers, then the entry address; after `ret` the stack
	// pointer must be 8 mod 16, as right after a call instruction
	uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
	void **sp = (void**)top;
	*--sp = 0;                     // fake return address of fibre_main (keeps the alignment rule)
	*--sp = (void*)fibre_main;
	for (int i = 0; i < 6; ++i) *--sp = 0;
	f.sp = sp;
	f.done = false;
}

void yield()
{
	Cta *c = cta;
	Fibre *f = c->cur;
	rb2emu_switch(&f->sp, c->sched_sp);
}

unsigned live_in_mask(unsigned w, unsigned mask)
{
	Cta *c = cta;
	unsigned n = 0;
	for (unsigned l = 0; l < 32; ++l) {
		const unsigned t = w * 32 + l;
		if ((mask >> l & 1) && t < c->nthreads && !c->fib[t].done) ++n;
	}
	return n;
}

struct StackPool {
	std::vector<char*> stacks;
	char *get(size_t i) {
		while (stacks.size() <= i) {
			void *p = mmap(0, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
			if (p == MAP_FAILED) { fprintf(stderr, "[cuda_emu] cannot map a fibre stack\n"); abort(); }
			stacks.push_back((char*)p);
		}
		return stacks[i];
	}
};
static thread_local StackPool pool;
static thread_local std::vector<unsigned char> dynbuf;

static thread_local const char *cur_kernel = 0;
static void on_segv(int sig)
{
	char msg[512];
	Cta *c = cta;
	int n = snprintf(msg, sizeof(msg), "[cuda_emu] signal %d in kernel %s, block (%u,%u,%u), thread %u\n", sig, cur_kernel ? cur_kernel : "(host code)",
	                 c ? c->bidx.x : 0, c ? c->bidx.y : 0, c ? c->bidx.z : 0, cur_threadIdx.x);
	if (write(2, msg, n) < 0) {}
	void *bt[32];
	int k = backtrace(bt, 32);
	backtrace_symbols_fd(bt, k, 2);
	_exit(139);
}
static void install_handler()
{
	static bool done = false;
	if (done) return;
	done = true;
	static char altstack[1 << 16];
	stack_t ss; ss.ss_sp = altstack; ss.ss_size = sizeof(altstack); ss.ss_flags = 0;
	sigaltstack(&ss, 0);
	struct sigaction sa; memset(&sa, 0, sizeof(sa));
	sa.sa_handler = on_segv; sa.sa_flags = SA_ONSTACK;
	sigaction(SIGSEGV, &sa, 0); sigaction(SIGBUS, &sa, 0);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body, const char *name)
{
	install_handler();
	cur_kernel = name;
	if (cta) { fprintf(stderr, "[cuda_emu] nested launch\n"); abort(); }
	const unsigned nt = block.x * block.y * block.z;
	if (nt == 0 || nt > 1024) { fprintf(stderr, "[cuda_emu] bad block size %u\n", nt); abort(); }
	if (smem > (227u << 10)) { fprintf(stderr, "[cuda_emu] %zu bytes of dynamic shared memory\n", smem); abort(); }
	Cta c;
	c.nthreads = nt; c.bdim = block; c.gdim = grid; c.body = &body;
	c.fib.resize(nt); c.warp.resize((nt + 31) / 32);
	if (dynbuf.size() < smem + 1024) dynbuf.resize(smem + 1024);
	c.dynsmem = (unsigned char*)(((uintptr_t)dynbuf.data() + 1023) & ~(uintptr_t)1023);
	for (unsigned t = 0; t < nt; ++t) { c.fib[t].stack = pool.get(t); c.fib[t].tid = t; }
	cta = &c;
	for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
		c.bidx = dim3(bx, by, bz);
		c.live = nt; c.progress = 0;
		for (int b = 0; b < 16; ++b) c.bars[b] = Bar();
		for (auto &w : c.warp) w.nmask = 0;
		memset(c.dynsmem, 0xCD, smem);
		for (unsigned t = 0; t < nt; ++t) fibre_init(c.fib[t]);
		unsigned idle = 0;
		while (c.live) {
			const unsigned before = c.progress;
			for (unsigned t = 0; t < nt; ++t) {
				Fibre &f = c.fib[t];
				if (f.done) continue;
				c.cur = &f;
				cur_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
				rb2emu_switch(&c.sched_sp, f.sp);
			}
			if (c.progress != before) idle = 0;
			else if (c.live) {
				// one more full round without any release: some barrier can never complete
				if (++idle > 4) { fprintf(stderr, "[cuda_emu] deadlock: %u threads of block (%u,%u,%u) wait at a barrier not every thread reaches\n", c.live, bx, by, bz); abort(); }
			}
		}
	}
	cta = 0;
	cur_kernel = 0;
}

} // namespace rb2emu
