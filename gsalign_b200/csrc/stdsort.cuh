// stdsort.cuh -- libstdc++'s std::sort, move for move, callable on the host and on the device.
//
// The reference sorts its block lists with std::sort under comparators that leave ties (src/GSAlign.cpp:17-27,
// src/ProcessCandidateAlignment.cpp:72-79), and the order libstdc++'s introsort leaves equal elements in is visible in the
// output (SURVEY.md hazards H5, H9).  Code that has to reproduce that order away from libstdc++ -- the block logic when it
// runs in a kernel -- therefore needs the SAME algorithm: introsort with a median-of-three pivot (first + 1, middle, last - 1
// moved to first), unguarded Hoare partition, depth limit 2 * floor(log2 n) with a heap-sort fallback, and a final insertion
// sort over stretches of 16 (bits/stl_algo.h: __introsort_loop, __unguarded_partition_pivot, __final_insertion_sort;
// bits/stl_heap.h).  The sequence of comparisons and moves only depends on comparison outcomes, so this restatement yields the
// same permutation as std::sort for any input (tests/test_boundary_cpu.py pins it against std::sort on tie-heavy data).
#pragma once
#if defined(__CUDACC__)
#define GSA_HD __host__ __device__
#else
#define GSA_HD
#endif

template <typename T> GSA_HD inline void gss_swap(T &a, T &b) { T t = a; a = b; b = t; }

// bits/stl_heap.h: __push_heap, __adjust_heap (value-based sift), __make_heap, __pop_heap, __sort_heap
template <typename T, typename Cmp>
GSA_HD inline void gss_push_heap(T *first, long hole, long top, T value, Cmp comp)
{
	long parent = (hole - 1) / 2;
	while (hole > top && comp(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
	first[hole] = value;
}

template <typename T, typename Cmp>
GSA_HD inline void gss_adjust_heap(T *first, long hole, long len, T value, Cmp comp)
{
	const long top = hole;
	long child = hole;
	while (child < (len - 1) / 2) {
		child = 2 * (child + 1);
		if (comp(first[child], first[child - 1])) child--;
		first[hole] = first[child];
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); first[hole] = first[child - 1]; hole = child - 1; }
	gss_push_heap(first, hole, top, value, comp);
}

template <typename T, typename Cmp>
GSA_HD inline void gss_make_heap(T *first, T *last, Cmp comp)
{
	const long len = last - first;
	if (len < 2) return;
	for (long parent = (len - 2) / 2;; parent--) {
		T value = first[parent];
		gss_adjust_heap(first, parent, len, value, comp);
		if (parent == 0) return;
	}
}

template <typename T, typename Cmp>
GSA_HD inline void gss_pop_heap(T *first, T *last, T *result, Cmp comp)
{
	T value = *result;
	*result = *first;
	gss_adjust_heap(first, 0L, (long)(last - first), value, comp);
}

// __partial_sort(first, last, last): __heap_select over an empty tail = __make_heap, then __sort_heap
template <typename T, typename Cmp>
GSA_HD inline void gss_heap_sort(T *first, T *last, Cmp comp)
{
	gss_make_heap(first, last, comp);
	while (last - first > 1) { --last; gss_pop_heap(first, last, last, comp); }
}

template <typename T, typename Cmp>
GSA_HD inline void gss_move_median_to_first(T *result, T *a, T *b, T *c, Cmp comp)
{
	if (comp(*a, *b)) {
		if (comp(*b, *c)) gss_swap(*result, *b);
		else if (comp(*a, *c)) gss_swap(*result, *c);
		else gss_swap(*result, *a);
	} else if (comp(*a, *c)) gss_swap(*result, *a);
	else if (comp(*b, *c)) gss_swap(*result, *c);
	else gss_swap(*result, *b);
}

template <typename T, typename Cmp>
GSA_HD inline T *gss_unguarded_partition(T *first, T *last, T *pivot, Cmp comp)
{
	for (;;) {
		while (comp(*first, *pivot)) ++first;
		--last;
		while (comp(*pivot, *last)) --last;
		if (!(first < last)) return first;
		gss_swap(*first, *last);
		++first;
	}
}

template <typename T, typename Cmp>
GSA_HD inline void gss_unguarded_linear_insert(T *last, Cmp comp)
{
	T val = *last;
	T *next = last - 1;
	while (comp(val, *next)) { *last = *next; last = next; --next; }
	*last = val;
}

template <typename T, typename Cmp>
GSA_HD inline void gss_insertion_sort(T *first, T *last, Cmp comp)
{
	if (first == last) return;
	for (T *i = first + 1; i != last; ++i) {
		if (comp(*i, *first)) {
			T val = *i;
			for (T *p = i; p != first; --p) *p = *(p - 1);   // move_backward(first, i, i + 1)
			*first = val;
		} else gss_unguarded_linear_insert(i, comp);
	}
}

// std::sort(first, last, comp).  The recursion of __introsort_loop (right part recursively, left part in the loop) is kept
// with an explicit stack: the order in which the parts are partitioned does not change any comparison, only when it happens.
template <typename T, typename Cmp>
GSA_HD inline void gsa_std_sort(T *first, T *last, Cmp comp)
{
	if (first == last) return;
	long n = last - first, lg = 0;
	while ((1L << (lg + 1)) <= n) lg++;                       // std::__lg
	struct Part { T *first, *last; long depth; };
	Part stack[128];
	int top = 0;
	stack[top++] = Part{first, last, 2 * lg};
	while (top > 0) {
		Part p = stack[--top];
		while (p.last - p.first > 16) {
			if (p.depth == 0) { gss_heap_sort(p.first, p.last, comp); break; }
			--p.depth;
			T *mid = p.first + (p.last - p.first) / 2;
			gss_move_median_to_first(p.first, p.first + 1, mid, p.last - 1, comp);
			T *cut = gss_unguarded_partition(p.first + 1, p.last, p.first, comp);
			if (top < 128) stack[top++] = Part{cut, p.last, p.depth};   // depth <= 2 lg n <= 126 parts pending at most
			p.last = cut;
		}
	}
	// __final_insertion_sort
	if (last - first > 16) {
		gss_insertion_sort(first, first + 16, comp);
		for (T *i = first + 16; i != last; ++i) gss_unguarded_linear_insert(i, comp);
	} else gss_insertion_sort(first, last, comp);
}
