__device__ float discount(float price, float rate) {
    return price * rate;
}
